// BandedBlockBandedMatrix * vector(s): descriptor, work partition and the two kernels.
//
//   bbb_walk_kernel     the fused band-walk kernel (hot path).  One CTA owns a tile of T in-block
//                       rows k in [k0,k0+T) and a segment of block rows [Ka,Kb).  It walks the
//                       block columns J = Ka-l .. Kb-1+u once; at step J it bulk-copies (TMA 1-D,
//                       cp.async.bulk + mbarrier ring) the *whole* contiguous P x (T+lam+mu) column
//                       chunk of block column J plus the matching x chunk into shared memory, and
//                       every thread (one output row each) gathers its l+u+1 slabs x lam+mu+1 band
//                       entries from it.  Partial sums for the l+u+1 block rows that block column
//                       J touches live in a register window that shifts by one per step; block row
//                       J-u is complete after step J and is written exactly once.
//                       => every byte of `data` is read from HBM once, in large contiguous pieces;
//                       y is written once; no atomics; summation order per entry is ascending
//                       global column, like the reference's per-block gbmv sweep.
//   bbb_generic_kernel  one thread per output row, direct global loads.  Takes every case the
//                       walk kernel does not (more than 8 block bands, P too large for shared
//                       memory, tiny blocks, empty bands).
//
// Reference behaviour being replaced: SURVEY.md 3.1 ([upstream] BlockArrays block loop + BLAS
// gbmv per block, operands from src/BandedBlockBandedMatrix.jl:503-545, bounds from
// src/AbstractBlockBandedMatrix.jl:90-104).
#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <new>

#include "bbm_internal.h"
#include "ptx.cuh"

namespace {

constexpr int kMaxNS = 8;  // block bands (l+u+1) the walk kernel keeps in registers

struct BbbItem {
    int32_t Ka, Kb;  // block-row segment [Ka,Kb)
    int32_t k0;      // first in-block row of the tile
    int32_t pad;
};

struct Partition {
    int tile = 0;
    int maxseg = 0;  // longest segment (block rows)
    int nitems = 0;
    BbbItem* d_items = nullptr;
    std::vector<BbbItem> items;  // host copy (halo-consumer counts of the fused multi-GPU path)
};

}  // namespace

struct bbm_bbb_desc {
    uint32_t magic = 0xBBBDE5C0u;
    bbm_handle_t h = nullptr;
    i64 Nr = 0, Nc = 0, l = 0, u = 0, lam = 0, mu = 0;
    i64 m = 0, n = 0, P = 0, maxr = 0;
    std::vector<i64> r, c, R, C;
    i64* dR = nullptr;
    i64* dC = nullptr;
    std::map<std::pair<int, int>, Partition> parts;  // (tile, target items) -> partition
    // Launch plans of the band-walk kernel: everything bbb_mul_impl derives from (descriptor, element type,
    // right-hand-side class, options, fused-halo geometry) -- tile, stages, partition, shared-memory size, the
    // kernel's address with its attribute already set -- so that a repeated product costs one cudaLaunchKernelExC
    // on the host (the 8-GPU step is ~30 us: a few tens of microseconds of host work per call would bound it).
    struct WalkPlan {
        i64 gen = -1;
        int elem = 0, nr = 0, peer = 0, own_J0 = 0, own_J1 = 0, force = 0;
        bool walk = false;
        int tile = 0, stages = 0, dbytes = 0, xbytes = 0, maxseg = 0, nitems = 0, pdl = 0, l2hint = 1;
        float l2keep = 0.0f;
        const BbbItem* d_items = nullptr;
        size_t smem = 0;
        int consumers[2] = {0, 0};
        const void* func = nullptr;
    };
    std::vector<WalkPlan> plans;
    // column tiles (block column, first column, count <= 256) of the staged product kernel with this matrix as C
    i64* d_mm_items = nullptr;
    i64 n_mm_items = 0;
    // block-row chunks of the pipelined host-vector product (sub-view descriptors, built on first use)
    struct HostChunk { i64 Ka, Kb, Jl, Jh; bbm_bbb_desc_t sub; };
    std::vector<HostChunk> chunks;
    std::vector<cudaEvent_t> chunk_ev;
};

namespace {

inline bool desc_valid(bbm_bbb_desc_t d) { return d && d->magic == 0xBBBDE5C0u && bbm_valid(d->h); }

template <typename T>
struct BbbParams {
    const T* data;
    const T* X;
    T* Y;
    i64 ldx, ldy;
    T alpha, beta;
    const i64* R;
    const i64* C;
    const BbbItem* items;
    i64 Nr, Nc, m;
    int l, u, lam, mu, w, P;
    int stages;
    int maxseg;
    int stage_data_bytes;  // per-stage bytes reserved for the matrix chunk
    int stage_x_bytes;     // per-stage bytes reserved for the x chunk
    int l2hint;
    int tri, unit;         // 0: whole matrix; 1 / 2: Upper- / LowerTriangular(A) (unit: ones on the diagonal)
    BbmPeerParams peer;    // fused halo push over NVLink peer memory (enabled == 0: plain kernel)
    int nrhs;              // right-hand sides (multi-vector kernel); keep new fields at the END: ptxas' code for the
                           // single-vector kernel proved sensitive to the parameter layout (5 % on cfg2)
    int pdl_early;         // != 0: release the dependent launch at the first step instead of the last one
    float l2keep;          // > 0: this fraction of the matrix lines is kept in L2 with priority ("bbb.l2keep_pct")
};

template <typename T> __device__ __forceinline__ T bbm_nan();
template <> __device__ __forceinline__ double bbm_nan<double>() { return __longlong_as_double(0x7ff8000000000000LL); }
template <> __device__ __forceinline__ float bbm_nan<float>() { return __int_as_float(0x7fc00000); }

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Wait for a flag written by another GPU.  timeout_clk == 0 waits for ever (what an NCCL receive would do);
// otherwise the wait gives up after that many SM clocks, latches the plan's error words (device block and the
// host-mapped word the next call checks) and returns false: the caller then poisons what it would have computed.
__device__ __forceinline__ bool peer_wait(const unsigned* flag, unsigned epoch, const BbmPeerParams& pe) {
    if ((int)(ld_acquire_sys(flag) - epoch) >= 0) return true;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
        __nanosleep(64);
        if (pe.timeout_clk > 0 && clock64() - t0 > pe.timeout_clk) {
            atomicExch(&pe.local->err, 1u);
            if (pe.err_host) { *reinterpret_cast<volatile unsigned*>(pe.err_host) = 1u; __threadfence_system(); }
            return false;
        }
    }
    return true;
}

// The pusher CTA of the fused multi-GPU step (blockIdx.x == 0, always in the first wave): store my boundary x
// blocks into the neighbours' halo buffers of this epoch's parity and publish the epoch.  The buffer was last
// read two epochs ago.  If the acknowledgement of that read never arrives, nothing is pushed and nothing is
// published (the neighbour's own wait then expires and it poisons its rows) -- never overwrite a buffer in use.
template <typename T>
__device__ __forceinline__ void peer_push(const BbmPeerParams& pe, const T* X, int tid, int nthreads) {
    __shared__ int push_ok[2];
    asm volatile("griddepcontrol.wait;" ::: "memory");   // x may come from the previous kernel in the stream
    if (tid == 0) {
#pragma unroll
        for (int sd = 0; sd < 2; ++sd) {
            push_ok[sd] = 1;
            if (pe.epoch >= 2 && pe.dst[sd] && pe.ack_wait[sd]) push_ok[sd] = peer_wait(&pe.local->ack_flag[sd], pe.epoch - 2, pe) ? 1 : 0;
        }
    }
    __syncthreads();
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
        if (!pe.dst[sd] || !push_ok[sd]) continue;
        T* dst = static_cast<T*>(pe.dst[sd]);
        const T* src = X + pe.src_off[sd];
        for (long long i = tid; i < pe.cnt[sd]; i += nthreads) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        // I am the RIGHT neighbour of my left neighbour, and vice versa
        if (pe.dst[0] && push_ok[0]) st_release_sys(&pe.remote[0]->data_flag[1], pe.epoch);
        if (pe.dst[1] && push_ok[1]) st_release_sys(&pe.remote[1]->data_flag[0], pe.epoch);
        // a halo nobody reads (no work item touches it) is acknowledged at once
#pragma unroll
        for (int sd = 0; sd < 2; ++sd)
            if (pe.ack_send[sd] && pe.consumers[sd] == 0) st_release_sys(&pe.remote[sd]->ack_flag[1 - sd], pe.epoch);
    }
}

// ------------------------------------------------------------------------------------------------
// band-walk kernel
// ------------------------------------------------------------------------------------------------
template <typename T>
struct StepGeom {
    int jlo, jhi, cJ;  // loaded in-block column range [jlo,jhi), block-column width
    i64 cbeg;          // first global column of block column J
};

template <typename T>
__device__ __forceinline__ StepGeom<T> step_geom(const i64* Ctab, int it, int k0, int tile, int lam, int mu) {
    StepGeom<T> g;
    g.cbeg = Ctab[it];
    g.cJ = static_cast<int>(Ctab[it + 1] - g.cbeg);
    g.jlo = max(0, k0 - lam);
    g.jhi = min(g.cJ, k0 + tile + mu);
    return g;
}

template <typename T, int NS, int W>
__global__ void __launch_bounds__(256) bbb_walk_kernel(const BbbParams<T> p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int tile = blockDim.x;
    const int w = (W > 0) ? W : p.w;
    const int P = p.P;
    const int S = p.stages;

    if (p.peer.enabled && blockIdx.x == 0) {
        if (blockIdx.y == 0) peer_push<T>(p.peer, p.X, tid, tile);
        return;
    }
    const BbbItem item = p.items[p.peer.enabled ? blockIdx.x - 1 : blockIdx.x];
    const int Ka = item.Ka, Kb = item.Kb, k0 = item.k0;
    const int Jfirst = Ka - p.l;
    const int nsteps = (Kb - Ka) + NS - 1;

    // shared memory carve-up
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                       // S barriers (<= 8)
    volatile int* poison = reinterpret_cast<volatile int*>(smem + 120);       // a halo never arrived: rows become NaN
    i64* Rtab = reinterpret_cast<i64*>(smem + 128);                           // maxseg+1
    i64* Ctab = Rtab + (p.maxseg + 1);                                        // maxseg+NS+1
    size_t tab_bytes = 128 + sizeof(i64) * (size_t)(2 * p.maxseg + NS + 2);
    tab_bytes = (tab_bytes + 127) & ~size_t(127);
    unsigned char* stage0 = smem + tab_bytes;
    const int stage_bytes = p.stage_data_bytes + p.stage_x_bytes;

    const T* __restrict__ Xq = p.X + (i64)blockIdx.y * p.ldx;
    T* __restrict__ Yq = p.Y + (i64)blockIdx.y * p.ldy;

    for (int i = tid; i <= Kb - Ka; i += tile) Rtab[i] = p.R[Ka + i];
    for (int i = tid; i <= nsteps; i += tile) {
        i64 J = (i64)Jfirst + i;
        J = J < 0 ? 0 : (J > p.Nc ? p.Nc : J);
        Ctab[i] = p.C[J];
    }
    if (tid == 0) {
        for (int s = 0; s < S; ++s) ptx::mbar_init(&bars[s], 1);
        ptx::fence_mbar_init();
        *poison = 0;
    }
    __syncthreads();
    // Programmatic dependent launch: everything above only touched the descriptor's tables (written at plan
    // time) and shared memory, so it may overlap the tail of the previous kernel in the stream; x, y and data
    // may have been produced by that kernel and are only touched after this point.
    asm volatile("griddepcontrol.wait;" ::: "memory");

    uint64_t pol_stream = 0, pol_keep = 0;
    if (tid == 0 && p.l2hint) {
        pol_stream = p.l2keep > 0.0f ? ptx::policy_keep_fraction(p.l2keep) : ptx::policy_evict_first();
        pol_keep = ptx::policy_evict_last();
    }

    // where step `it` takes its x block from: x itself, or (fused multi-GPU path) the halo buffer a
    // neighbour filled; both indexed by local column
    auto xsource = [&](int it, int* side) -> const T* {
        *side = -1;
        if (p.peer.enabled) {
            const int Jl = Jfirst + it;
            *side = Jl < p.peer.own_J0 ? 0 : (Jl >= p.peer.own_J1 ? 1 : -1);
            if (*side >= 0) return static_cast<const T*>(p.peer.halo[*side]) - p.peer.halo_col0[*side];
        }
        return Xq;
    };

    // producer: one thread issues the bulk copies of a step and arms its barrier
    auto issue = [&](int it) {
        const int st = it % S;
        const StepGeom<T> g = step_geom<T>(Ctab, it, k0, tile, p.lam, p.mu);
        if (g.jlo >= g.jhi) {
            ptx::mbar_arrive(&bars[st]);  // nothing to load for this step
            return;
        }
        unsigned char* sd = stage0 + (size_t)st * stage_bytes;
        unsigned char* sx = sd + p.stage_data_bytes;
        const uintptr_t d0 = reinterpret_cast<uintptr_t>(p.data + (g.cbeg + g.jlo) * (i64)P);
        const uintptr_t d1 = reinterpret_cast<uintptr_t>(p.data + (g.cbeg + g.jhi) * (i64)P);
        int halo_side = -1;
        const T* xsrc = xsource(it, &halo_side);
        const uintptr_t x0 = reinterpret_cast<uintptr_t>(xsrc + g.cbeg + g.jlo);
        const uintptr_t x1 = reinterpret_cast<uintptr_t>(xsrc + g.cbeg + g.jhi);
        const uintptr_t dg0 = d0 & ~uintptr_t(15), dg1 = (d1 + 15) & ~uintptr_t(15);
        const uintptr_t xg0 = x0 & ~uintptr_t(15), xg1 = (x1 + 15) & ~uintptr_t(15);
        const uint32_t dbytes = static_cast<uint32_t>(dg1 - dg0);
        const uint32_t xbytes = static_cast<uint32_t>(xg1 - xg0);
        ptx::mbar_arrive_expect_tx(&bars[st], dbytes + xbytes);
        if (p.l2hint) ptx::bulk_g2s_hint(sd, reinterpret_cast<const void*>(dg0), dbytes, &bars[st], pol_stream);
        else ptx::bulk_g2s(sd, reinterpret_cast<const void*>(dg0), dbytes, &bars[st]);
        if (halo_side >= 0) {
            // a halo block column: its x block is stored by the neighbour's pusher CTA -- wait for its epoch
            // flag, then order the (generic-proxy) acquire before the async-proxy bulk read
            if (!peer_wait(&p.peer.local->data_flag[halo_side], p.peer.epoch, p.peer)) *poison = 1;
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        if (p.l2hint) ptx::bulk_g2s_hint(sx, reinterpret_cast<const void*>(xg0), xbytes, &bars[st], pol_keep);
        else ptx::bulk_g2s(sx, reinterpret_cast<const void*>(xg0), xbytes, &bars[st]);
    };

    if (tid == 0) {
        for (int it = 0; it < S - 1 && it < nsteps; ++it) issue(it);
    }

    T acc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[s] = T(0);

    const int k = k0 + tid;
    const T alpha = p.alpha, beta = p.beta;

    for (int it = 0; it < nsteps; ++it) {
        // every thread has finished reading the stage that step it+S-1 will overwrite
        __syncthreads();
        if (it == (p.pdl_early ? 0 : nsteps - 1)) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // successor may set up
        if (tid == 0 && it + S - 1 < nsteps) issue(it + S - 1);

        const int st = it % S;
        const int J = Jfirst + it;
        const StepGeom<T> g = step_geom<T>(Ctab, it, k0, tile, p.lam, p.mu);
        ptx::mbar_wait(&bars[st], (uint32_t)((it / S) & 1));

        if (g.jlo < g.jhi) {
            const unsigned char* sd = stage0 + (size_t)st * stage_bytes;
            const unsigned char* sx = sd + p.stage_data_bytes;
            const uintptr_t d0 = reinterpret_cast<uintptr_t>(p.data + (g.cbeg + g.jlo) * (i64)P);
            int side_unused;
            const uintptr_t x0 = reinterpret_cast<uintptr_t>(xsource(it, &side_unused) + g.cbeg + g.jlo);
            // Ds[(j-jlo)*P + row] = data[row, cbeg+j];  Xs[j-jlo] = x[cbeg+j]
            const T* Ds = reinterpret_cast<const T*>(sd + (d0 & 15));
            const T* Xs = reinterpret_cast<const T*>(sx + (x0 & 15));
            const int cJ = g.cJ, jlo = g.jlo;
            if (W > 0) {
                // the w x-values of this row are shared by all NS slabs: fetch them once
                constexpr int WW = W > 0 ? W : 1;
                T xr[WW];
                int off[WW];
#pragma unroll
                for (int ii = 0; ii < WW; ++ii) {
                    const int j = k + p.mu - ii;
                    const bool ok = j >= 0 && j < cJ;
                    off[ii] = ok ? (j - jlo) * P + ii : -1;
                    xr[ii] = ok ? Xs[j - jlo] : T(0);
                }
                // which of the NS block rows this step feeds are live for this thread
                bool act[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const int K = J + s - p.u;
                    const bool in = K >= Ka && K < Kb;
                    const int rK = in ? static_cast<int>(Rtab[K - Ka + 1] - Rtab[K - Ka]) : 0;
                    act[s] = k < rK;
                }
                // column-major over the sub-band (ascending j per accumulator, as the reference), the NS
                // accumulators interleaved: branch-free straight-line code with NS independent FMA chains
#pragma unroll
                for (int ii = WW - 1; ii >= 0; --ii) {
                    const int o = off[ii] >= 0 ? off[ii] : 0;   // any in-chunk address; its value is discarded
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        T dv = Ds[o + s * WW];
                        bool keep = act[s] && off[ii] >= 0;
                        if (p.tri) {
                            // slab s is block (J+s-u, J): above the diagonal iff s < u; inside the diagonal block
                            // entry (k,j) has j - k = mu - ii
                            const bool up = p.tri == 1;
                            keep = keep && (s == p.u ? (up ? ii <= p.mu : ii >= p.mu) : (up ? s < p.u : s > p.u));
                            if (p.unit && s == p.u && ii == p.mu) dv = T(1);
                        }
                        const T r = fma(dv, xr[ii], acc[s]);
                        acc[s] = keep ? r : acc[s];   // never lets an unused slot (NaN) in
                    }
                }
            } else {
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const int K = J + s - p.u;
                    if (K >= Ka && K < Kb) {  // block-uniform
                        const int rK = static_cast<int>(Rtab[K - Ka + 1] - Rtab[K - Ka]);
                        if (k < rK) {
                            for (int ii = w - 1; ii >= 0; --ii) {
                                const int j = k + p.mu - ii;
                                if (j < 0 || j >= cJ) continue;
                                T dv = Ds[(j - jlo) * P + s * w + ii];
                                if (p.tri) {
                                    const bool up = p.tri == 1;
                                    if (!(s == p.u ? (up ? ii <= p.mu : ii >= p.mu) : (up ? s < p.u : s > p.u))) continue;
                                    if (p.unit && s == p.u && ii == p.mu) dv = T(1);
                                }
                                acc[s] = fma(dv, Xs[j - jlo], acc[s]);
                            }
                        }
                    }
                }
            }
        }

        // block row J-u has now received every block column it overlaps: write it out once
        const int Kdone = J - p.u;
        if (Kdone >= Ka && Kdone < Kb) {
            const i64 r0 = Rtab[Kdone - Ka];
            const int rK = static_cast<int>(Rtab[Kdone - Ka + 1] - r0);
            if (k < rK) {
                T y = alpha * acc[0];
                if (beta != T(0)) y = fma(beta, Yq[r0 + k], y);
                if (p.peer.enabled && *poison) y = bbm_nan<T>();
                Yq[r0 + k] = y;
            }
        }
#pragma unroll
        for (int s = 0; s + 1 < NS; ++s) acc[s] = acc[s + 1];
        acc[NS - 1] = T(0);
    }

    if (p.peer.enabled && blockIdx.y == 0) {
        // tell the neighbour whose data I read that its next push may overwrite my halo
        const BbmPeerParams& pe = p.peer;
        const bool used[2] = {Jfirst < pe.own_J0, Jfirst + nsteps - 1 >= pe.own_J1};
        if (tid == 0) {
#pragma unroll
            for (int sd = 0; sd < 2; ++sd) {
                if (!used[sd] || !pe.ack_send[sd]) continue;
                __threadfence();
                const unsigned c = atomicAdd(&pe.local->count[sd], 1u);
                if (c + 1 == (unsigned)pe.consumers[sd]) {
                    pe.local->count[sd] = 0;
                    __threadfence_system();
                    st_release_sys(&pe.remote[sd]->ack_flag[1 - sd], pe.epoch);
                }
            }
        }
    }
}

// Multi-vector variant (A*X): NR right-hand sides per CTA -- the matrix chunk of a step is staged once and
// multiplied with NR x chunks, so A is streamed once per group of NR columns instead of once per column.
// Kept apart from the single-vector kernel above: folding NR into it cost the headline path 5-10 %
// (different unrolling / register allocation), measured on cfg2 and cfg3.
template <typename T, int NS, int W, int NR>
__global__ void __launch_bounds__(256) bbb_walk_mr_kernel(const BbbParams<T> p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int tile = blockDim.x;
    const int w = (W > 0) ? W : p.w;
    const int P = p.P;
    const int S = p.stages;

    if (p.peer.enabled && blockIdx.x == 0) {
        if (blockIdx.y == 0) peer_push<T>(p.peer, p.X, tid, tile);
        return;
    }
    const BbbItem item = p.items[p.peer.enabled ? blockIdx.x - 1 : blockIdx.x];
    const int Ka = item.Ka, Kb = item.Kb, k0 = item.k0;
    const int Jfirst = Ka - p.l;
    const int nsteps = (Kb - Ka) + NS - 1;

    // shared memory carve-up
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                       // S barriers (<= 8)
    volatile int* poison = reinterpret_cast<volatile int*>(smem + 120);       // a halo never arrived: rows become NaN
    i64* Rtab = reinterpret_cast<i64*>(smem + 128);                           // maxseg+1
    i64* Ctab = Rtab + (p.maxseg + 1);                                        // maxseg+NS+1
    size_t tab_bytes = 128 + sizeof(i64) * (size_t)(2 * p.maxseg + NS + 2);
    tab_bytes = (tab_bytes + 127) & ~size_t(127);
    unsigned char* stage0 = smem + tab_bytes;
    const int stage_bytes = p.stage_data_bytes + p.stage_x_bytes;

    const int q0 = blockIdx.y * NR;
    const int nq = NR == 1 ? 1 : min(NR, p.nrhs - q0);   // live right-hand sides of this CTA
    const T* __restrict__ Xq = p.X + (i64)q0 * p.ldx;
    T* __restrict__ Yq = p.Y + (i64)q0 * p.ldy;
    const int xone = p.stage_x_bytes / NR;                // bytes reserved per x chunk

    for (int i = tid; i <= Kb - Ka; i += tile) Rtab[i] = p.R[Ka + i];
    for (int i = tid; i <= nsteps; i += tile) {
        i64 J = (i64)Jfirst + i;
        J = J < 0 ? 0 : (J > p.Nc ? p.Nc : J);
        Ctab[i] = p.C[J];
    }
    if (tid == 0) {
        for (int s = 0; s < S; ++s) ptx::mbar_init(&bars[s], 1);
        ptx::fence_mbar_init();
        *poison = 0;
    }
    __syncthreads();
    // Programmatic dependent launch: everything above only touched the descriptor's tables (written at plan
    // time) and shared memory, so it may overlap the tail of the previous kernel in the stream; x, y and data
    // may have been produced by that kernel and are only touched after this point.
    asm volatile("griddepcontrol.wait;" ::: "memory");

    uint64_t pol_stream = 0, pol_keep = 0;
    if (tid == 0 && p.l2hint) {
        pol_stream = ptx::policy_evict_first();
        pol_keep = ptx::policy_evict_last();
    }

    // where step `it` takes its x block from: x itself, or (fused multi-GPU path) the halo buffer a
    // neighbour filled; both indexed by local column
    auto xsource = [&](int it, int* side) -> const T* {
        *side = -1;
        if (p.peer.enabled) {
            const int Jl = Jfirst + it;
            *side = Jl < p.peer.own_J0 ? 0 : (Jl >= p.peer.own_J1 ? 1 : -1);
            if (*side >= 0) return static_cast<const T*>(p.peer.halo[*side]) - p.peer.halo_col0[*side];
        }
        return Xq;
    };

    // producer: one thread issues the bulk copies of a step and arms its barrier
    auto issue = [&](int it) {
        const int st = it % S;
        const StepGeom<T> g = step_geom<T>(Ctab, it, k0, tile, p.lam, p.mu);
        if (g.jlo >= g.jhi) {
            ptx::mbar_arrive(&bars[st]);  // nothing to load for this step
            return;
        }
        unsigned char* sd = stage0 + (size_t)st * stage_bytes;
        unsigned char* sx = sd + p.stage_data_bytes;
        const uintptr_t d0 = reinterpret_cast<uintptr_t>(p.data + (g.cbeg + g.jlo) * (i64)P);
        const uintptr_t d1 = reinterpret_cast<uintptr_t>(p.data + (g.cbeg + g.jhi) * (i64)P);
        int halo_side = -1;
        const T* xsrc = xsource(it, &halo_side);
        const uintptr_t x0 = reinterpret_cast<uintptr_t>(xsrc + g.cbeg + g.jlo);
        const uintptr_t x1 = reinterpret_cast<uintptr_t>(xsrc + g.cbeg + g.jhi);
        const uintptr_t dg0 = d0 & ~uintptr_t(15), dg1 = (d1 + 15) & ~uintptr_t(15);
        const uintptr_t xg0 = x0 & ~uintptr_t(15), xg1 = (x1 + 15) & ~uintptr_t(15);
        const uint32_t dbytes = static_cast<uint32_t>(dg1 - dg0);
        const uint32_t xbytes = static_cast<uint32_t>(xg1 - xg0);
        uint32_t xtotal = xbytes;
        if (NR > 1) {
            xtotal = 0;
#pragma unroll
            for (int q = 0; q < NR; ++q) {
                if (q >= nq) break;
                const uintptr_t sh = (uintptr_t)q * (uintptr_t)p.ldx * sizeof(T);
                xtotal += static_cast<uint32_t>((((x1 + sh) + 15) & ~uintptr_t(15)) - ((x0 + sh) & ~uintptr_t(15)));
            }
        }
        ptx::mbar_arrive_expect_tx(&bars[st], dbytes + xtotal);
        if (p.l2hint) ptx::bulk_g2s_hint(sd, reinterpret_cast<const void*>(dg0), dbytes, &bars[st], pol_stream);
        else ptx::bulk_g2s(sd, reinterpret_cast<const void*>(dg0), dbytes, &bars[st]);
        if (halo_side >= 0) {
            // a halo block column: its x block is stored by the neighbour's pusher CTA -- wait for its epoch
            // flag, then order the (generic-proxy) acquire before the async-proxy bulk read
            if (!peer_wait(&p.peer.local->data_flag[halo_side], p.peer.epoch, p.peer)) *poison = 1;
            asm volatile("fence.proxy.async;" ::: "memory");
        }
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            if (q >= nq) break;
            // columns of X may start at different 16-byte phases: every column gets its own aligned window
            const uintptr_t sh = (uintptr_t)q * (uintptr_t)p.ldx * sizeof(T);
            const uintptr_t g0 = (x0 + sh) & ~uintptr_t(15), g1 = ((x1 + sh) + 15) & ~uintptr_t(15);
            const uint32_t nb = static_cast<uint32_t>(g1 - g0);
            if (p.l2hint) ptx::bulk_g2s_hint(sx + q * xone, reinterpret_cast<const void*>(g0), nb, &bars[st], pol_keep);
            else ptx::bulk_g2s(sx + q * xone, reinterpret_cast<const void*>(g0), nb, &bars[st]);
        }
    };

    if (tid == 0) {
        for (int it = 0; it < S - 1 && it < nsteps; ++it) issue(it);
    }

    T acc[NS][NR];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int q = 0; q < NR; ++q) acc[s][q] = T(0);

    const int k = k0 + tid;
    const T alpha = p.alpha, beta = p.beta;

    for (int it = 0; it < nsteps; ++it) {
        // every thread has finished reading the stage that step it+S-1 will overwrite
        __syncthreads();
        if (it == (p.pdl_early ? 0 : nsteps - 1)) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // successor may set up
        if (tid == 0 && it + S - 1 < nsteps) issue(it + S - 1);

        const int st = it % S;
        const int J = Jfirst + it;
        const StepGeom<T> g = step_geom<T>(Ctab, it, k0, tile, p.lam, p.mu);
        ptx::mbar_wait(&bars[st], (uint32_t)((it / S) & 1));

        if (g.jlo < g.jhi) {
            const unsigned char* sd = stage0 + (size_t)st * stage_bytes;
            const unsigned char* sx = sd + p.stage_data_bytes;
            const uintptr_t d0 = reinterpret_cast<uintptr_t>(p.data + (g.cbeg + g.jlo) * (i64)P);
            int side_unused;
            const uintptr_t x0 = reinterpret_cast<uintptr_t>(xsource(it, &side_unused) + g.cbeg + g.jlo);
            // Ds[(j-jlo)*P + row] = data[row, cbeg+j];  Xs[j-jlo] = x[cbeg+j]
            const T* Ds = reinterpret_cast<const T*>(sd + (d0 & 15));
            const T* Xs[NR];
#pragma unroll
            for (int q = 0; q < NR; ++q) Xs[q] = reinterpret_cast<const T*>(sx + q * xone + ((x0 + (uintptr_t)q * (uintptr_t)p.ldx * sizeof(T)) & 15));
            const int cJ = g.cJ, jlo = g.jlo;
            if (W > 0) {
                // the w x-values of this row are shared by all NS slabs: fetch them once
                constexpr int WW = W > 0 ? W : 1;
                T xr[WW][NR];
                int off[WW];
#pragma unroll
                for (int ii = 0; ii < WW; ++ii) {
                    const int j = k + p.mu - ii;
                    const bool ok = j >= 0 && j < cJ;
                    off[ii] = ok ? (j - jlo) * P + ii : -1;
#pragma unroll
                    for (int q = 0; q < NR; ++q) xr[ii][q] = (ok && q < nq) ? Xs[q][j - jlo] : T(0);
                }
                // which of the NS block rows this step feeds are live for this thread
                bool act[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const int K = J + s - p.u;
                    const bool in = K >= Ka && K < Kb;
                    const int rK = in ? static_cast<int>(Rtab[K - Ka + 1] - Rtab[K - Ka]) : 0;
                    act[s] = k < rK;
                }
                // column-major over the sub-band (ascending j per accumulator, as the reference), the NS
                // accumulators interleaved: branch-free straight-line code with NS independent FMA chains
#pragma unroll
                for (int ii = WW - 1; ii >= 0; --ii) {
                    const int o = off[ii] >= 0 ? off[ii] : 0;   // any in-chunk address; its value is discarded
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        T dv = Ds[o + s * WW];
                        bool keep = act[s] && off[ii] >= 0;
                        if (p.tri) {
                            // slab s is block (J+s-u, J): above the diagonal iff s < u; inside the diagonal block
                            // entry (k,j) has j - k = mu - ii
                            const bool up = p.tri == 1;
                            keep = keep && (s == p.u ? (up ? ii <= p.mu : ii >= p.mu) : (up ? s < p.u : s > p.u));
                            if (p.unit && s == p.u && ii == p.mu) dv = T(1);
                        }
#pragma unroll
                        for (int q = 0; q < NR; ++q) {
                            const T r = fma(dv, xr[ii][q], acc[s][q]);
                            acc[s][q] = keep ? r : acc[s][q];   // never lets an unused slot (NaN) in
                        }
                    }
                }
            } else {
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const int K = J + s - p.u;
                    if (K >= Ka && K < Kb) {  // block-uniform
                        const int rK = static_cast<int>(Rtab[K - Ka + 1] - Rtab[K - Ka]);
                        if (k < rK) {
                            for (int ii = w - 1; ii >= 0; --ii) {
                                const int j = k + p.mu - ii;
                                if (j < 0 || j >= cJ) continue;
                                T dv = Ds[(j - jlo) * P + s * w + ii];
                                if (p.tri) {
                                    const bool up = p.tri == 1;
                                    if (!(s == p.u ? (up ? ii <= p.mu : ii >= p.mu) : (up ? s < p.u : s > p.u))) continue;
                                    if (p.unit && s == p.u && ii == p.mu) dv = T(1);
                                }
#pragma unroll
                                for (int q = 0; q < NR; ++q)
                                    if (q < nq) acc[s][q] = fma(dv, Xs[q][j - jlo], acc[s][q]);
                            }
                        }
                    }
                }
            }
        }

        // block row J-u has now received every block column it overlaps: write it out once
        const int Kdone = J - p.u;
        if (Kdone >= Ka && Kdone < Kb) {
            const i64 r0 = Rtab[Kdone - Ka];
            const int rK = static_cast<int>(Rtab[Kdone - Ka + 1] - r0);
            if (k < rK) {
#pragma unroll
                for (int q = 0; q < NR; ++q) {
                    if (q >= nq) break;
                    T* yp = Yq + (i64)q * p.ldy + r0 + k;
                    T y = alpha * acc[0][q];
                    if (beta != T(0)) y = fma(beta, *yp, y);
                    if (p.peer.enabled && *poison) y = bbm_nan<T>();
                    *yp = y;
                }
            }
        }
#pragma unroll
        for (int s = 0; s + 1 < NS; ++s)
#pragma unroll
            for (int q = 0; q < NR; ++q) acc[s][q] = acc[s + 1][q];
#pragma unroll
        for (int q = 0; q < NR; ++q) acc[NS - 1][q] = T(0);
    }

    if (p.peer.enabled && blockIdx.y == 0) {
        // tell the neighbour whose data I read that its next push may overwrite my halo
        const BbmPeerParams& pe = p.peer;
        const bool used[2] = {Jfirst < pe.own_J0, Jfirst + nsteps - 1 >= pe.own_J1};
        if (tid == 0) {
#pragma unroll
            for (int sd = 0; sd < 2; ++sd) {
                if (!used[sd] || !pe.ack_send[sd]) continue;
                __threadfence();
                const unsigned c = atomicAdd(&pe.local->count[sd], 1u);
                if (c + 1 == (unsigned)pe.consumers[sd]) {
                    pe.local->count[sd] = 0;
                    __threadfence_system();
                    st_release_sys(&pe.remote[sd]->ack_flag[1 - sd], pe.epoch);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// generic kernel: one thread per output row
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bbb_generic_kernel(const BbbParams<T> p) {
    const T* __restrict__ Xq = p.X + (i64)blockIdx.y * p.ldx;
    T* __restrict__ Yq = p.Y + (i64)blockIdx.y * p.ldy;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 row = (i64)blockIdx.x * blockDim.x + threadIdx.x; row < p.m; row += stride) {
        // K = last block with R[K] <= row (zero-size blocks share a start; take the non-empty one)
        i64 lo = 0, hi = p.Nr;  // invariant: R[lo] <= row < R[hi]
        while (hi - lo > 1) {
            const i64 mid = (lo + hi) >> 1;
            if (p.R[mid] <= row) lo = mid; else hi = mid;
        }
        const i64 K = lo;
        const i64 k = row - p.R[K];
        T acc = T(0);
        if (p.P > 0) {
            const i64 Jlo = max((i64)0, K - p.l), Jhi = min(p.Nc - 1, K + p.u);
            for (i64 J = Jlo; J <= Jhi; ++J) {
                const i64 cbeg = p.C[J];
                const i64 cJ = p.C[J + 1] - cbeg;
                const T* slab = p.data + (i64)(p.u + K - J) * p.w;
                for (int ii = p.w - 1; ii >= 0; --ii) {
                    const i64 j = k + p.mu - ii;
                    if (j < 0 || j >= cJ) continue;
                    if (p.tri) {
                        const bool up = p.tri == 1;
                        if (!(J == K ? (up ? ii <= p.mu : ii >= p.mu) : (up ? J > K : J < K))) continue;
                    }
                    T dv = slab[ii + (i64)p.P * (cbeg + j)];
                    if (p.tri && p.unit && J == K && ii == p.mu) dv = T(1);
                    acc = fma(dv, Xq[cbeg + j], acc);
                }
            }
        }
        T y = p.alpha * acc;
        if (p.beta != T(0)) y = fma(p.beta, Yq[row], y);
        Yq[row] = y;
    }
}

// ------------------------------------------------------------------------------------------------
// host side: partition + dispatch
// ------------------------------------------------------------------------------------------------
size_t walk_tab_bytes(int maxseg, int NS) {
    size_t b = 128 + sizeof(i64) * (size_t)(2 * maxseg + NS + 2);
    return (b + 127) & ~size_t(127);
}

template <typename T>
void walk_stage_bytes(int tile, int w, i64 P, int* dbytes, int* xbytes) {
    size_t d = (size_t)(tile + w - 1) * (size_t)P * sizeof(T) + 32;
    size_t x = (size_t)(tile + w - 1) * sizeof(T) + 32;
    *dbytes = (int)((d + 15) & ~size_t(15));
    *xbytes = (int)((x + 15) & ~size_t(15));
}

// Split the (tile, block-row) work into at most `slots` items of about equal weight, `slots` being a
// whole number of waves of resident CTAs (measured on B200: 1 wave beats 1.5 or 3.03 waves by
// 20-30%, a partial last wave cannot saturate HBM).  A tile t covers in-block rows [t*T, (t+1)*T);
// only block rows with r[K] > t*T take part, in maximal runs; every run gets a share of the slots
// proportional to its weight and is cut into that many segments of equal weight.
// lead_steps > 0 (fused multi-GPU path): the items that start at block row 0 first wait for the left
// neighbour's halo push (a few microseconds of NVLink latency and fences); they get that many steps less
// work so that they still finish with the rest of the wave.
int32_t build_partition(bbm_bbb_desc_t d, int tile, int slots, int maxseg_cap, int lead_steps, Partition* out) {
    bbm_handle_t h = d->h;
    const i64 Nr = d->Nr;
    // per-step cost that does not shrink with the tile's height.  Measured on cfg3 (block sizes 1..4000): a step
    // costs nearly the same whether its tile is full or almost empty (it is latency-, not byte-bound), so
    // the items are balanced by step count with the row count as a tie-breaker: 0.61 -> 0.96 of HBM peak (F32)
    const i64 ovh = std::max<i64>(0, bbm_opt(h, "bbb.ovh", 4096));
    struct Run { int t; i64 Ka, Kb; double wt; int nseg; };
    std::vector<Run> runs;
    const i64 ntiles = (d->maxr + tile - 1) / tile;
    double total = 0;
    auto weight = [&](i64 K, i64 k0) {
        const double wgt = (double)(std::min<i64>(tile, d->r[K] - k0) + ovh);
        return K == 0 ? wgt * (1 + lead_steps) : wgt;
    };
    for (i64 t = 0; t < ntiles; ++t) {
        const i64 k0 = t * tile;
        i64 K = 0;
        while (K < Nr) {
            if (d->r[K] <= k0) { ++K; continue; }
            Run run{(int)t, K, K, 0.0, 1};
            while (K < Nr && d->r[K] > k0) { run.wt += weight(K, k0); ++K; }
            run.Kb = K;
            total += run.wt;
            runs.push_back(run);
        }
    }
    // segments per run: proportional share of the slots, at least what the table size demands, at
    // most one per block row; leftovers go to the runs whose segments are heaviest
    i64 used = 0;
    for (Run& run : runs) {
        const i64 len = run.Kb - run.Ka;
        i64 n = (i64)((double)slots * run.wt / std::max(total, 1.0));
        n = std::max<i64>(n, (len + maxseg_cap - 1) / maxseg_cap);
        n = std::min<i64>(std::max<i64>(n, 1), len);
        run.nseg = (int)n;
        used += n;
    }
    while (used < slots) {
        Run* best = nullptr;
        for (Run& run : runs)
            if (run.nseg < run.Kb - run.Ka && (!best || run.wt / run.nseg > best->wt / best->nseg)) best = &run;
        if (!best) break;
        ++best->nseg;
        ++used;
    }
    std::vector<BbbItem> items;
    int maxseg = 1;
    for (const Run& run : runs) {
        const i64 k0 = (i64)run.t * tile;
        i64 Ka = run.Ka;
        double done = 0;
        for (int sidx = 0; sidx < run.nseg && Ka < run.Kb; ++sidx) {
            const double goal = run.wt * (sidx + 1) / run.nseg;
            i64 Kb = Ka;
            const i64 must_leave = run.nseg - 1 - sidx;  // one block row at least for every later segment
            while (Kb < run.Kb - must_leave && (Kb == Ka || done + 0.5 * weight(Kb, k0) <= goal) && (Kb - Ka) < maxseg_cap) {
                done += weight(Kb, k0);
                ++Kb;
            }
            if (sidx == run.nseg - 1) {
                // last segment takes the rest (split further only if the table size forces it)
                while (run.Kb - Ka > maxseg_cap) {
                    items.push_back(BbbItem{(int32_t)Ka, (int32_t)(Ka + maxseg_cap), (int32_t)k0, 0});
                    maxseg = std::max(maxseg, maxseg_cap);
                    Ka += maxseg_cap;
                }
                Kb = run.Kb;
            }
            items.push_back(BbbItem{(int32_t)Ka, (int32_t)Kb, (int32_t)k0, 0});
            maxseg = std::max<int>(maxseg, (int)(Kb - Ka));
            Ka = Kb;
        }
    }
    // adjacent tiles of the same segment read adjacent memory: order by (Ka, tile)
    std::stable_sort(items.begin(), items.end(), [](const BbbItem& a, const BbbItem& b) {
        return a.Ka != b.Ka ? a.Ka < b.Ka : a.k0 < b.k0;
    });
    out->tile = tile;
    out->maxseg = maxseg;
    out->nitems = (int)items.size();
    out->items = items;
    out->d_items = nullptr;
    if (!items.empty()) {
        cudaError_t e = cudaMalloc(&out->d_items, items.size() * sizeof(BbbItem));
        if (e != cudaSuccess) { cudaGetLastError(); return bbm_fail(h, BBM_ERR_ALLOC, "cudaMalloc(items): %s", cudaGetErrorString(e)); }
        BBM_CUDA(h, cudaMemcpy(out->d_items, items.data(), items.size() * sizeof(BbbItem), cudaMemcpyHostToDevice));
    }
    return BBM_OK;
}

// address of the band-walk kernel instantiation for (NS block bands, w sub-bands, nr right-hand sides per CTA)
template <typename T, int NS>
const void* walk_kernel_w(int w, int nr) {
    if (nr == 4) {
        if (w == 3) return reinterpret_cast<const void*>(&bbb_walk_mr_kernel<T, NS, 3, 4>);
        if (w == 5) return reinterpret_cast<const void*>(&bbb_walk_mr_kernel<T, NS, 5, 4>);
        return reinterpret_cast<const void*>(&bbb_walk_mr_kernel<T, NS, 0, 4>);
    }
    if (w == 3) return reinterpret_cast<const void*>(&bbb_walk_kernel<T, NS, 3>);
    if (w == 5) return reinterpret_cast<const void*>(&bbb_walk_kernel<T, NS, 5>);
    return reinterpret_cast<const void*>(&bbb_walk_kernel<T, NS, 0>);
}

template <typename T>
const void* walk_kernel(int NS, int w, int nr) {
    switch (NS) {
        case 1: return walk_kernel_w<T, 1>(w, nr);
        case 2: return walk_kernel_w<T, 2>(w, nr);
        case 3: return walk_kernel_w<T, 3>(w, nr);
        case 4: return walk_kernel_w<T, 4>(w, nr);
        case 5: return walk_kernel_w<T, 5>(w, nr);
        case 6: return walk_kernel_w<T, 6>(w, nr);
        case 7: return walk_kernel_w<T, 7>(w, nr);
        case 8: return walk_kernel_w<T, 8>(w, nr);
        default: return nullptr;
    }
}

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the (context, function) pair, not to a launch plan: plans of
// several descriptors share one kernel instantiation, and a later plan with a smaller footprint must not lower the
// limit under an earlier, cached one ("invalid argument" at its next launch).  The limit is therefore only ever raised.
cudaError_t ensure_dyn_smem(int device, const void* func, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> have;
    std::lock_guard<std::mutex> lk(mu);
    size_t& cur = have[std::make_pair(device, func)];
    if (bytes <= cur) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}

// Builds the launch plan of the band-walk kernel for (descriptor, T, nr, options, fused-halo geometry); plan->walk
// stays false when the generic kernel has to take the product.
template <typename T>
int32_t build_walk_plan(bbm_handle_t h, bbm_bbb_desc_t d, int nr, bbm_bbb_desc::WalkPlan* plan) {
    const i64 NS = d->l + d->u + 1;
    const int w = (int)(d->lam + d->mu + 1);
    const i64 force = bbm_opt(h, "bbb.kernel", 0);
    plan->walk = false;
    plan->l2hint = (int)bbm_opt(h, "bbb.l2hint", 1);
    // "bbb.l2keep_pct" (default 0 = stream the matrix): for back-to-back products with a matrix not much larger than L2
    // (an iterative solver on one rank's shard), keep that percentage of its lines resident across products
    plan->l2keep = (float)std::min<i64>(100, std::max<i64>(0, bbm_opt(h, "bbb.l2keep_pct", 0))) / 100.0f;
    bool walk = d->P > 0 && NS >= 1 && NS <= kMaxNS && force != 2;
    // tiny blocks leave the 32..256-row tiles mostly idle: use the per-row kernel
    if (walk && force != 1 && d->m < 16 * d->Nr && !h->peer) walk = false;
    if (!walk) return BBM_OK;
    const void* func = walk_kernel<T>((int)NS, w, nr);
    if (!func) return BBM_OK;

    int tile = 0, stages = 0, dbytes = 0, xbytes = 0;
    const i64 forced_tile = bbm_opt(h, "bbb.tile", 0);
    if (forced_tile && (forced_tile < 32 || forced_tile > 256 || forced_tile % 32))
        return bbm_fail(h, BBM_ERR_ARG, "bbb.tile must be a multiple of 32 in [32,256]");
    const size_t budget = std::min<size_t>(h->smem_optin, 200 * 1024);
    const size_t soft = (228 * 1024 - 2 * 1024) / 2 - 256;  // two CTAs per SM (1 KB reserved per CTA)
    const int cands[4] = {forced_tile ? (int)forced_tile : 256, 128, 64, 32};
    const int ncand = forced_tile ? 1 : 4;
    const size_t tabs = walk_tab_bytes(512, (int)NS);
    auto choose = [&](int want_stages) {
        tile = 0; stages = 0;
        // pass 0: >= 3 stages within the two-CTA budget; pass 1: >= 2 stages within the opt-in limit
        for (int pass = 0; pass < 2 && !tile; ++pass) {
            const size_t limit = pass == 0 ? soft : budget;
            const int min_stages = pass == 0 ? 3 : 2;
            for (int ci = 0; ci < ncand && !tile; ++ci) {
                const int t = cands[ci];
                if (!forced_tile && t > 32 && t / 2 >= d->maxr) continue;  // tile twice the largest block: mostly idle
                walk_stage_bytes<T>(t, w, d->P, &dbytes, &xbytes);
                const size_t per = (size_t)dbytes + (size_t)nr * xbytes;
                if (limit <= tabs) continue;
                const int s = (int)std::min<size_t>(want_stages, (limit - tabs) / per);
                if (s >= min_stages) { tile = t; stages = s; }
            }
        }
    };
    // Stage count.  Measured on B200 (2-D Laplacian, 4096-row blocks): with long items (221 steps each at 4096 block
    // rows) 4 stages x 2 CTAs per SM saturate HBM (231 us; 3 stages: 245 us; 5-6 stages 5 % slower); with short items
    // (28 steps each at 512 block rows, one rank's share of the 8-GPU run) 3 stages let a THIRD CTA fit per SM and the
    // step drops from 34.4 to 30.8 us (27.2 us with the dependent launch on top): the kernel is then bound by per-CTA
    // latencies (prologue, barrier, tail), which a third resident CTA hides.  "bbb.stages" overrides.
    const i64 opt_stages = bbm_opt(h, "bbb.stages", 0);
    choose((int)std::min<i64>(8, std::max<i64>(2, opt_stages > 0 ? opt_stages : 4)));
    if (tile && opt_stages <= 0 && stages > 3) {
        i64 units = 0;
        for (i64 K = 0; K < d->Nr; ++K) units += (d->r[K] + tile - 1) / tile;
        if (units < (i64)64 * 2 * h->num_sms) choose(3);
    }
    if (!tile) return BBM_OK;  // P too large for shared memory staging

    walk_stage_bytes<T>(tile, w, d->P, &dbytes, &xbytes);
    const int maxseg_cap = (int)std::min<i64>(512, std::max<i64>(1, bbm_opt(h, "bbb.maxseg", 512)));
    // slots = whole waves of resident CTAs (table size taken at its cap: a few KB either way).  "bbb.ctas_per_sm"
    // overrides the occupancy (programmatic dependent launch wants a free slot per SM for the successor's CTAs).
    int occ = 1;
    {
        const size_t smem_cap = walk_tab_bytes(maxseg_cap, (int)NS) + (size_t)stages * ((size_t)dbytes + (size_t)nr * xbytes);
        cudaError_t e = ensure_dyn_smem(h->device, func, smem_cap);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, func, tile, smem_cap);
        if (e != cudaSuccess) { cudaGetLastError(); occ = 1; }
        occ = std::max(1, occ);
        const i64 cps = bbm_opt(h, "bbb.ctas_per_sm", 0);
        if (cps > 0) occ = (int)std::min<i64>(occ, cps);
    }
    const i64 ips = bbm_opt(h, "bbb.items_per_sm", 0);
    const i64 waves = std::max<i64>(1, bbm_opt(h, "bbb.waves", 1));
    // the fused multi-GPU path adds one pusher CTA: leave it a slot of the wave
    const int target_items = (int)(ips > 0 ? h->num_sms * ips : h->num_sms * occ * waves) - (h->peer ? 1 : 0);
    const int lead_steps = (h->peer && h->peer->has_halo[0]) ? (int)std::max<i64>(0, bbm_opt(h, "dist.lead_steps", 0)) : 0;
    auto key = std::make_pair(tile * 64 + lead_steps + 65536 * (int)bbm_opt(h, "bbb.ovh", 4096), target_items * 1024 + maxseg_cap);
    auto it = d->parts.find(key);
    if (it == d->parts.end()) {
        Partition part;
        int32_t rc = build_partition(d, tile, target_items, maxseg_cap, lead_steps, &part);
        if (rc != BBM_OK) return rc;
        it = d->parts.emplace(key, part).first;
    }
    const Partition& part = it->second;
    if (part.nitems == 0) return BBM_OK;
    plan->consumers[0] = plan->consumers[1] = 0;
    if (h->peer) {
        for (const BbbItem& bi : part.items) {
            if ((i64)bi.Ka - d->l < h->peer->own_J0) ++plan->consumers[0];
            if ((i64)bi.Kb - 1 + d->u >= h->peer->own_J1) ++plan->consumers[1];
        }
    }
    plan->tile = tile; plan->stages = stages; plan->dbytes = dbytes; plan->xbytes = xbytes;
    plan->maxseg = part.maxseg; plan->nitems = part.nitems; plan->d_items = part.d_items;
    plan->smem = walk_tab_bytes(part.maxseg, (int)NS) + (size_t)stages * ((size_t)dbytes + (size_t)nr * xbytes);
    plan->func = func;
    // "bbb.pdl": 0 plain launch; 1 programmatic dependent launch released at the last step; 2 (default) released at
    // the first step: the successor's CTAs are placed as soon as slots free up, read their tables and arm their
    // barriers while this kernel still streams, and block in griddepcontrol.wait before touching x, y or data.
    // Measured (back-to-back products, one event at either end): 34.4 -> 31.2 us at 512 block rows, 27.2 us together
    // with the third CTA per SM; 231.1 -> 228.2 us at 4096 block rows.
    plan->pdl = (int)bbm_opt(h, "bbb.pdl", 2);
    cudaError_t e = ensure_dyn_smem(h->device, func, plan->smem);
    if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "cudaFuncSetAttribute(bbb_walk_kernel): %s", cudaGetErrorString(e));
    plan->walk = true;
    return BBM_OK;
}

template <typename T>
int32_t bbb_mul_impl(bbm_handle_t h, bbm_bbb_desc_t d, T alpha, const T* d_data, const T* d_X, i64 ldx, i64 nrhs,
                     T beta, T* d_Y, i64 ldy, int tri = 0, int unit = 0) {
    BBM_REQUIRE_HANDLE(h);
    if (!desc_valid(d)) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BandedBlockBandedMatrix descriptor");
    if (nrhs < 0) return bbm_fail(h, BBM_ERR_ARG, "nrhs < 0");
    if (nrhs > 0 && (ldx < std::max<i64>(1, d->n) || ldy < std::max<i64>(1, d->m)))
        return bbm_fail(h, BBM_ERR_DIM, "DimensionMismatch: A is %lld x %lld, ldx=%lld, ldy=%lld", (long long)d->m,
                        (long long)d->n, (long long)ldx, (long long)ldy);
    if (d->m == 0 || nrhs == 0) return BBM_OK;
    if (!d_Y || (d->n > 0 && !d_X) || (d->P > 0 && d->n > 0 && !d_data))
        return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    if (nrhs > 65535) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "nrhs > 65535");
    BBM_CUDA(h, cudaSetDevice(h->device));

    BbbParams<T> p{};
    p.data = d_data; p.X = d_X; p.Y = d_Y; p.ldx = ldx; p.ldy = ldy; p.alpha = alpha; p.beta = beta;
    p.R = d->dR; p.C = d->dC; p.Nr = d->Nr; p.Nc = d->Nc; p.m = d->m;
    p.l = (int)d->l; p.u = (int)d->u; p.lam = (int)d->lam; p.mu = (int)d->mu;
    p.w = (int)(d->lam + d->mu + 1); p.P = (int)d->P;
    p.tri = tri; p.unit = unit; p.nrhs = (int)nrhs;
    if (tri && (d->Nr != d->Nc || d->r != d->c))  // hasmatchingblocks, src/triblockbanded.jl:12,20
        return bbm_fail(h, BBM_ERR_UNSUPPORTED, "triangular product needs matching square diagonal blocks");
    const bool aligned = (reinterpret_cast<uintptr_t>(d_data) % sizeof(T) == 0) && (reinterpret_cast<uintptr_t>(d_X) % sizeof(T) == 0);
    if (!aligned) return bbm_fail(h, BBM_ERR_ARG, "device pointers must be aligned to the element size");
    if (h->peer && nrhs != 1) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "fused halo push needs the band-walk kernel and one right-hand side");

    const int nr = (nrhs >= 2 && !h->peer) ? 4 : 1;  // right-hand sides per CTA of the band-walk kernel
    bbm_bbb_desc::WalkPlan* plan = nullptr;
    for (auto& pl : d->plans)
        if (pl.gen == h->opt_gen && pl.elem == (int)sizeof(T) && pl.nr == nr && pl.peer == (h->peer ? 1 : 0) &&
            (!h->peer || (pl.own_J0 == h->peer->own_J0 && pl.own_J1 == h->peer->own_J1))) { plan = &pl; break; }
    if (!plan) {
        bbm_bbb_desc::WalkPlan np;
        np.gen = h->opt_gen; np.elem = (int)sizeof(T); np.nr = nr; np.peer = h->peer ? 1 : 0;
        if (h->peer) { np.own_J0 = h->peer->own_J0; np.own_J1 = h->peer->own_J1; }
        int32_t rc = build_walk_plan<T>(h, d, nr, &np);
        if (rc != BBM_OK) return rc;
        if (d->plans.size() >= 16) d->plans.erase(d->plans.begin());   // stale option generations
        d->plans.push_back(np);
        plan = &d->plans.back();
    }
    p.l2hint = plan->l2hint;
    if (h->peer && !plan->walk) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "fused halo push needs the band-walk kernel");

    if (plan->walk) {
        p.items = plan->d_items;
        if (h->peer) {
            p.peer = *h->peer;
            p.peer.consumers[0] = plan->consumers[0];
            p.peer.consumers[1] = plan->consumers[1];
        }
        p.stages = plan->stages;
        p.maxseg = plan->maxseg;
        p.stage_data_bytes = plan->dbytes;
        p.stage_x_bytes = nr * plan->xbytes;
        p.pdl_early = plan->pdl == 2 ? 1 : 0;
        p.l2keep = plan->l2keep;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)plan->nitems + (h->peer ? 1u : 0u), (unsigned)((nrhs + nr - 1) / nr), 1);
        cfg.blockDim = dim3((unsigned)plan->tile);
        cfg.dynamicSmemBytes = plan->smem;
        cfg.stream = h->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = plan->pdl ? 1 : 0;
        void* args[1] = {&p};
        cudaError_t e = cudaLaunchKernelExC(&cfg, plan->func, args);
        if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "bbb_walk_kernel launch: %s", cudaGetErrorString(e));
        h->launches += 1;
        return BBM_OK;
    }
    {
        const int threads = 256;
        const i64 blocks = std::min<i64>((d->m + threads - 1) / threads, (i64)h->num_sms * 32);
        dim3 grid((unsigned)std::max<i64>(1, blocks), (unsigned)nrhs, 1);
        bbb_generic_kernel<T><<<grid, threads, 0, h->stream>>>(p);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "bbb_generic_kernel launch: %s", cudaGetErrorString(e));
        h->launches += 1;
    }
    return BBM_OK;
}

// ------------------------------------------------------------------------------------------------
// transposed product  Y <- alpha*A^T*X + beta*Y  (`mul!(y, A', x)`; layouts of src/adjtransblockbanded.jl:2-7)
// In band storage column c of `data` holds exactly the entries of column c of A, i.e. of row c of A^T: output
// entry c is the dot product of that contiguous column (P values) with gathered x -- no accumulation across
// columns, no atomics.  One thread per column; for a fixed storage row the threads of a warp read x at
// consecutive rows (coalesced) and `data` at stride P (every byte of the touched lines is used by the warp's
// next storage rows: L1 absorbs the stride).  Summation order: ascending storage row = ascending (K, k).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bbb_mul_t_kernel(const BbbParams<T> p, i64 n) {
    const T* __restrict__ Xq = p.X + (i64)blockIdx.y * p.ldx;
    T* __restrict__ Yq = p.Y + (i64)blockIdx.y * p.ldy;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const int NS = p.w > 0 ? p.P / p.w : 0;
    for (i64 col = (i64)blockIdx.x * blockDim.x + threadIdx.x; col < n; col += stride) {
        i64 lo = 0, hi = p.Nc;  // J = last block column with C[J] <= col
        while (hi - lo > 1) {
            const i64 mid = (lo + hi) >> 1;
            if (p.C[mid] <= col) lo = mid; else hi = mid;
        }
        const i64 J = lo, j = col - p.C[J];
        const T* __restrict__ dcol = p.data + (i64)p.P * col;
        T acc = T(0);
        for (int s = 0; s < NS; ++s) {
            const i64 K = J + s - p.u;
            if (K < 0 || K >= p.Nr) continue;
            const i64 r0 = p.R[K], rK = p.R[K + 1] - r0;
            for (int ii = 0; ii < p.w; ++ii) {
                const i64 k = j - p.mu + ii;
                if (k >= 0 && k < rK) acc = fma(dcol[s * p.w + ii], Xq[r0 + k], acc);
            }
        }
        T y = p.alpha * acc;
        if (p.beta != T(0)) y = fma(p.beta, Yq[col], y);
        Yq[col] = y;
    }
}

template <typename T>
int32_t bbb_mul_t_impl(bbm_handle_t h, bbm_bbb_desc_t d, T alpha, const T* d_data, const T* d_X, i64 ldx, i64 nrhs, T beta,
                       T* d_Y, i64 ldy) {
    BBM_REQUIRE_HANDLE(h);
    if (!desc_valid(d)) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BandedBlockBandedMatrix descriptor");
    if (nrhs < 0) return bbm_fail(h, BBM_ERR_ARG, "nrhs < 0");
    if (nrhs > 0 && (ldx < std::max<i64>(1, d->m) || ldy < std::max<i64>(1, d->n)))
        return bbm_fail(h, BBM_ERR_DIM, "DimensionMismatch: A' is %lld x %lld, ldx=%lld, ldy=%lld", (long long)d->n,
                        (long long)d->m, (long long)ldx, (long long)ldy);
    if (d->n == 0 || nrhs == 0) return BBM_OK;
    if (!d_Y || (d->m > 0 && !d_X) || (d->P > 0 && !d_data)) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    if (nrhs > 65535) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "nrhs > 65535");
    BBM_CUDA(h, cudaSetDevice(h->device));
    BbbParams<T> p{};
    p.data = d_data; p.X = d_X; p.Y = d_Y; p.ldx = ldx; p.ldy = ldy; p.alpha = alpha; p.beta = beta;
    p.R = d->dR; p.C = d->dC; p.Nr = d->Nr; p.Nc = d->Nc; p.m = d->m;
    p.l = (int)d->l; p.u = (int)d->u; p.lam = (int)d->lam; p.mu = (int)d->mu;
    p.w = (int)std::max<i64>(0, d->lam + d->mu + 1); p.P = (int)d->P;
    const i64 blocks = std::min<i64>((d->n + 255) / 256, (i64)h->num_sms * 32);
    bbb_mul_t_kernel<T><<<dim3((unsigned)std::max<i64>(1, blocks), (unsigned)nrhs), 256, 0, h->stream>>>(p, d->n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "bbb_mul_t_kernel launch: %s", cudaGetErrorString(e));
    h->launches += 1;
    return BBM_OK;
}

template <typename T>
int32_t bbb_mul_host_impl(bbm_handle_t h, bbm_bbb_desc_t d, T alpha, const T* d_data, const T* h_X, i64 ldx, i64 nrhs,
                          T beta, T* h_Y, i64 ldy) {
    BBM_REQUIRE_HANDLE(h);
    if (!desc_valid(d)) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BandedBlockBandedMatrix descriptor");
    if (nrhs < 0) return bbm_fail(h, BBM_ERR_ARG, "nrhs < 0");
    if (nrhs > 0 && (ldx < std::max<i64>(1, d->n) || ldy < std::max<i64>(1, d->m)))
        return bbm_fail(h, BBM_ERR_DIM, "DimensionMismatch: A is %lld x %lld, ldx=%lld, ldy=%lld", (long long)d->m,
                        (long long)d->n, (long long)ldx, (long long)ldy);
    if (d->m == 0 || nrhs == 0) return BBM_OK;
    if (!h_Y || (d->n > 0 && !h_X)) return bbm_fail(h, BBM_ERR_ARG, "null host pointer");
    BBM_CUDA(h, cudaSetDevice(h->device));
    const i64 n1 = std::max<i64>(1, d->n), m1 = std::max<i64>(1, d->m);
    int32_t rc = bbm_ws_reserve(h, &h->ws_x, &h->ws_x_bytes, (size_t)(n1 * nrhs) * sizeof(T) + 64);
    if (rc != BBM_OK) return rc;
    rc = bbm_ws_reserve(h, &h->ws_y, &h->ws_y_bytes, (size_t)(m1 * nrhs) * sizeof(T) + 64);
    if (rc != BBM_OK) return rc;
    T* dX = static_cast<T*>(h->ws_x);
    T* dY = static_cast<T*>(h->ws_y);

    // Large single-vector products are PCIe-bound (x up, y down): pipeline them over block-row chunks on
    // three streams -- upload the x columns chunk i needs, multiply chunk i (a sub-view: block rows
    // [Ka,Kb) are banded-block-banded with shifted bandwidths over a column range of the same storage,
    // src/BandedBlockBandedMatrix.jl:416-420,450-456), download y of chunk i -- so that the two copy
    // directions and the kernel overlap and the call costs about max(H2D, D2H) instead of their sum.
    const i64 nchunk_opt = bbm_opt(h, "bbb.host_chunks", 8);
    const bool pipelined = nrhs == 1 && nchunk_opt > 1 && d->P > 0 && d->Nr >= 2 * nchunk_opt &&
                           (size_t)(d->m + d->n) * sizeof(T) >= (size_t)(8u << 20);
    if (pipelined) {
        if (d->chunks.empty()) {
            const int nch = (int)std::min<i64>(nchunk_opt, 32);
            std::vector<i64> kb(nch + 1);
            if (bbm_bbb_partition_rows(d->Nr, d->r.data(), nch, kb.data()) != BBM_OK) return bbm_fail(h, BBM_ERR_ARG, "bad block sizes");
            for (int i = 0; i < nch; ++i) {
                if (kb[i + 1] <= kb[i]) continue;
                bbm_bbb_desc::HostChunk ch{kb[i], kb[i + 1], 0, 0, nullptr};
                ch.Jl = std::min<i64>(std::max<i64>(ch.Ka - d->l, 0), d->Nc);
                ch.Jh = std::min<i64>(std::max<i64>(ch.Kb - 1 + d->u + 1, ch.Jl), d->Nc);
                const i64 sft = ch.Ka - ch.Jl;
                rc = bbm_bbb_desc_create(h, ch.Kb - ch.Ka, ch.Jh - ch.Jl, d->r.data() + ch.Ka, d->c.data() + ch.Jl, d->l - sft,
                                         d->u + sft, d->lam, d->mu, &ch.sub);
                if (rc != BBM_OK) return rc;
                d->chunks.push_back(ch);
            }
            d->chunk_ev.resize(2 * d->chunks.size() + 1, nullptr);
            for (auto& e : d->chunk_ev) BBM_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        if (!h->copy_in) {
            BBM_CUDA(h, cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking));
            BBM_CUDA(h, cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking));
        }
        cudaEvent_t ev0 = d->chunk_ev.back();
        BBM_CUDA(h, cudaEventRecord(ev0, h->stream));            // earlier work on the handle's stream
        BBM_CUDA(h, cudaStreamWaitEvent(h->copy_in, ev0, 0));
        BBM_CUDA(h, cudaStreamWaitEvent(h->copy_out, ev0, 0));
        i64 up = 0;  // x columns [0, up) are on the device
        for (size_t i = 0; i < d->chunks.size(); ++i) {
            const auto& ch = d->chunks[i];
            const i64 need = d->C[ch.Jh];
            if (need > up) {
                BBM_CUDA(h, cudaMemcpyAsync(dX + up, h_X + up, (size_t)(need - up) * sizeof(T), cudaMemcpyHostToDevice, h->copy_in));
                up = need;
            }
            const i64 r0 = d->R[ch.Ka], nr = d->R[ch.Kb] - r0;
            if (beta != T(0) && nr > 0)
                BBM_CUDA(h, cudaMemcpyAsync(dY + r0, h_Y + r0, (size_t)nr * sizeof(T), cudaMemcpyHostToDevice, h->copy_in));
            BBM_CUDA(h, cudaEventRecord(d->chunk_ev[2 * i], h->copy_in));
            BBM_CUDA(h, cudaStreamWaitEvent(h->stream, d->chunk_ev[2 * i], 0));
            const i64 c0 = d->C[ch.Jl];
            rc = bbb_mul_impl<T>(h, ch.sub, alpha, d_data + d->P * c0, dX + c0, n1, 1, beta, dY + r0, m1);
            if (rc != BBM_OK) return rc;
            BBM_CUDA(h, cudaEventRecord(d->chunk_ev[2 * i + 1], h->stream));
            BBM_CUDA(h, cudaStreamWaitEvent(h->copy_out, d->chunk_ev[2 * i + 1], 0));
            if (nr > 0)
                BBM_CUDA(h, cudaMemcpyAsync(h_Y + r0, dY + r0, (size_t)nr * sizeof(T), cudaMemcpyDeviceToHost, h->copy_out));
        }
        BBM_CUDA(h, cudaStreamSynchronize(h->copy_out));
        BBM_CUDA(h, cudaStreamSynchronize(h->copy_in));
        BBM_CUDA(h, cudaStreamSynchronize(h->stream));
        return BBM_OK;
    }

    if (d->n > 0)
        BBM_CUDA(h, cudaMemcpy2DAsync(dX, n1 * sizeof(T), h_X, ldx * sizeof(T), d->n * sizeof(T), nrhs, cudaMemcpyHostToDevice, h->stream));
    if (beta != T(0))
        BBM_CUDA(h, cudaMemcpy2DAsync(dY, m1 * sizeof(T), h_Y, ldy * sizeof(T), d->m * sizeof(T), nrhs, cudaMemcpyHostToDevice, h->stream));
    rc = bbb_mul_impl<T>(h, d, alpha, d_data, dX, n1, nrhs, beta, dY, m1);
    if (rc != BBM_OK) return rc;
    BBM_CUDA(h, cudaMemcpy2DAsync(h_Y, ldy * sizeof(T), dY, m1 * sizeof(T), d->m * sizeof(T), nrhs, cudaMemcpyDeviceToHost, h->stream));
    BBM_CUDA(h, cudaStreamSynchronize(h->stream));
    return BBM_OK;
}


// ------------------------------------------------------------------------------------------------
// BandedBlockBandedMatrix x BandedBlockBandedMatrix -> BandedBlockBandedMatrix
// ([upstream] _block_muladd! with one banded x banded gbmm per block triple; result structure
// src/linalg.jl:50-71).  One thread per stored entry of C: it gathers the (<= (lA+uA+1) blocks) x
// (<= min(wA,wB) inner indices) products of its row of A and column of B in ascending (N, i) order.
// Every entry of C's storage is written exactly once (unused slots get beta*old, or 0 for beta == 0, like
// the reference's _fill_lmul!), so there are no atomics; HBM-bound on C with A and B served by L1/L2.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct BbbMmParams {
    const T* A; const T* B; T* C;
    const i64 *RA, *CA, *CB;   // prefix sums: rows of A (= rows of C), cols of A (= rows of B), cols of B (= cols of C)
    i64 NrA, NcA, NcB, nC;
    int lA, uA, lamA, muA, wA, PA;
    int lB, uB, lamB, muB, wB, PB;
    int lC, uC, lamC, muC, wC, PC;
    T alpha, beta;
};

template <typename T>
__global__ void __launch_bounds__(256) bbb_matmul_kernel(const BbbMmParams<T> p) {
    const i64 total = (i64)p.PC * p.nC;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const i64 col = e / p.PC;
        const int rho = (int)(e - col * p.PC);
        const int sC = rho / p.wC, ii = rho - sC * p.wC;
        // J = block column of `col` (last block with CB[J] <= col; empty blocks share a start)
        i64 lo = 0, hi = p.NcB;
        while (hi - lo > 1) {
            const i64 mid = (lo + hi) >> 1;
            if (p.CB[mid] <= col) lo = mid; else hi = mid;
        }
        const i64 J = lo, j = col - p.CB[J];
        const i64 K = J + sC - p.uC;
        const i64 k = j - p.muC + ii;
        T acc = T(0);
        const bool live = K >= 0 && K < p.NrA && k >= 0 && k < p.RA[K + 1] - p.RA[K];
        if (live && p.PA > 0 && p.PB > 0) {
            const i64 Nlo = max(max(K - p.lA, J - p.uB), (i64)0), Nhi = min(min(K + p.uA, J + p.lB), p.NcA - 1);
            for (i64 N = Nlo; N <= Nhi; ++N) {
                const i64 cN = p.CA[N + 1] - p.CA[N];
                const i64 ilo = max(max(k - p.lamA, j - p.muB), (i64)0), ihi = min(min(k + p.muA, j + p.lamB), cN - 1);
                const T* __restrict__ a = p.A + (i64)(p.uA + K - N) * p.wA + p.muA + k;        // + (-i) + PA*(CA[N]+i)
                const T* __restrict__ b = p.B + (i64)(p.uB + N - J) * p.wB + p.muB - j + (i64)p.PB * col;  // + i
                for (i64 i = ilo; i <= ihi; ++i) acc = fma(a[-i + (i64)p.PA * (p.CA[N] + i)], b[i], acc);
            }
        }
        T v = p.alpha * acc;
        if (p.beta != T(0)) v = fma(p.beta, p.C[e], v);
        p.C[e] = v;
    }
}

// Column-per-thread variant: column (J,j) of C is A times column (J,j) of B.  The thread reads its P_B
// contiguous entries of B once and, for each, the P_A contiguous entries of one column of A (neighbouring
// threads read neighbouring columns of A: L1 hits); the P_C partial sums live in shared memory ([rho][thread],
// padded, conflict-free both ways) and the CTA's 256 x P_C result slab -- contiguous in C's storage -- is written
// back coalesced.  The storage row of C an (A entry, B entry) pair lands in depends only on the two storage rows,
// not on the column, so no index search happens in the inner loops.  Ascending (N, i) order per entry.
template <typename T>
__global__ void __launch_bounds__(256) bbb_matmul_col_kernel(const BbbMmParams<T> p) {
    extern __shared__ __align__(16) unsigned char mmc_smem[];
    T* acc = reinterpret_cast<T*>(mmc_smem);  // acc[rho * 257 + t]
    const int tid = threadIdx.x;
    const int nsA = p.PA / max(p.wA, 1), nsB = p.PB / max(p.wB, 1);
    for (i64 col0 = (i64)blockIdx.x * 256; col0 < p.nC; col0 += (i64)gridDim.x * 256) {
        const i64 col = col0 + tid;
        for (int rho = 0; rho < p.PC; ++rho) acc[rho * 257 + tid] = T(0);
        if (col < p.nC && p.PA > 0 && p.PB > 0) {
            i64 lo = 0, hi = p.NcB;
            while (hi - lo > 1) {
                const i64 mid = (lo + hi) >> 1;
                if (p.CB[mid] <= col) lo = mid; else hi = mid;
            }
            const i64 J = lo, j = col - p.CB[J];
            const T* __restrict__ bcol = p.B + (i64)p.PB * col;
            for (int sB = 0; sB < nsB; ++sB) {
                const i64 N = J + sB - p.uB;
                if (N < 0 || N >= p.NcA) continue;
                const i64 cN = p.CA[N + 1] - p.CA[N];
                for (int iB = 0; iB < p.wB; ++iB) {
                    const i64 i = j - p.muB + iB;
                    if (i < 0 || i >= cN) continue;
                    const T b = bcol[sB * p.wB + iB];
                    const T* __restrict__ acol = p.A + (i64)p.PA * (p.CA[N] + i);
                    for (int sA = 0; sA < nsA; ++sA) {
                        const i64 K = N + sA - p.uA;
                        if (K < 0 || K >= p.NrA) continue;
                        const i64 rK = p.RA[K + 1] - p.RA[K];
                        const int sC = p.uC - p.uA - p.uB + sA + sB;          // = uC + K - J
                        for (int iA = 0; iA < p.wA; ++iA) {
                            const i64 k = i - p.muA + iA;
                            if (k < 0 || k >= rK) continue;
                            const int rho = sC * p.wC + (p.muC - p.muA - p.muB + iA + iB);   // iiC = muC + k - j
                            acc[rho * 257 + tid] = fma(acol[sA * p.wA + iA], b, acc[rho * 257 + tid]);
                        }
                    }
                }
            }
        }
        __syncthreads();
        const i64 ncol = min((i64)256, p.nC - col0);
        T* __restrict__ cslab = p.C + (i64)p.PC * col0;
        for (i64 idx = tid; idx < ncol * p.PC; idx += 256) {
            const int t = (int)(idx / p.PC), rho = (int)(idx - (i64)t * p.PC);
            T v = p.alpha * acc[rho * 257 + t];
            if (p.beta != T(0)) v = fma(p.beta, cslab[idx], v);
            cslab[idx] = v;
        }
        __syncthreads();
    }
}

// Compile-time bandwidths for the column kernel (NSA/NSB block bands, WA/WB sub-block bands of A and B): the
// four product loops unroll completely, the NS x NW = (NSA+NSB-1) x (WA+WB-1) partial sums of a column of C
// live in registers instead of being read-modified-written in shared memory, and block sizes are looked up once
// per column.  Shared memory only turns the 256 x P_C slab for the coalesced write-back.  Same ascending (N, i)
// summation order as the generic kernel.
template <typename T, int NSA, int WA, int NSB, int WB>
__global__ void __launch_bounds__(256) bbb_matmul_col_fixed_kernel(const BbbMmParams<T> p) {
    extern __shared__ __align__(16) unsigned char mmc_smem[];
    T* acc = reinterpret_cast<T*>(mmc_smem);  // acc[rho * 257 + t]
    constexpr int NS = NSA + NSB - 1, NW = WA + WB - 1;
    const int tid = threadIdx.x;
    const int s0 = p.uC - p.uA - p.uB, i0 = p.muC - p.muA - p.muB;   // >= 0 (BandError otherwise)
    for (i64 col0 = (i64)blockIdx.x * 256; col0 < p.nC; col0 += (i64)gridDim.x * 256) {
        const i64 col = col0 + tid;
        T r[NS][NW];
#pragma unroll
        for (int a = 0; a < NS; ++a)
#pragma unroll
            for (int b = 0; b < NW; ++b) r[a][b] = T(0);
        if (col < p.nC) {
            i64 lo = 0, hi = p.NcB;
            while (hi - lo > 1) {
                const i64 mid = (lo + hi) >> 1;
                if (p.CB[mid] <= col) lo = mid; else hi = mid;
            }
            const i64 J = lo, j = col - p.CB[J];
            i64 rK[NS];   // sizes of the block rows K = J - uA - uB + a of A (0 outside the matrix)
#pragma unroll
            for (int a = 0; a < NS; ++a) {
                const i64 K = J - p.uA - p.uB + a;
                rK[a] = (K >= 0 && K < p.NrA) ? p.RA[K + 1] - p.RA[K] : 0;
            }
            const T* __restrict__ bcol = p.B + (i64)p.PB * col;
#pragma unroll
            for (int sB = 0; sB < NSB; ++sB) {
                const i64 N = J + sB - p.uB;
                if (N < 0 || N >= p.NcA) continue;
                const i64 cN0 = p.CA[N], cN = p.CA[N + 1] - cN0;
#pragma unroll
                for (int iB = 0; iB < WB; ++iB) {
                    const i64 i = j - p.muB + iB;
                    if (i < 0 || i >= cN) continue;
                    const T b = bcol[sB * WB + iB];
                    const T* __restrict__ acol = p.A + (i64)(NSA * WA) * (cN0 + i);
#pragma unroll
                    for (int sA = 0; sA < NSA; ++sA) {
#pragma unroll
                        for (int iA = 0; iA < WA; ++iA) {
                            const i64 k = i - p.muA + iA;
                            if (k >= 0 && k < rK[sA + sB]) r[sA + sB][iA + iB] = fma(acol[sA * WA + iA], b, r[sA + sB][iA + iB]);
                        }
                    }
                }
            }
        }
        for (int rho = 0; rho < p.PC; ++rho) acc[rho * 257 + tid] = T(0);
#pragma unroll
        for (int a = 0; a < NS; ++a)
#pragma unroll
            for (int b = 0; b < NW; ++b) acc[((s0 + a) * p.wC + i0 + b) * 257 + tid] = r[a][b];
        __syncthreads();
        const i64 ncol = min((i64)256, p.nC - col0);
        T* __restrict__ cslab = p.C + (i64)p.PC * col0;
        for (i64 idx = tid; idx < ncol * p.PC; idx += 256) {
            const int t = (int)(idx / p.PC), rho = (int)(idx - (i64)t * p.PC);
            T v = p.alpha * acc[rho * 257 + t];
            if (p.beta != T(0)) v = fma(p.beta, cslab[idx], v);
            cslab[idx] = v;
        }
        __syncthreads();
    }
}


// Staged variant of the fixed-shape column kernel.  The column kernels above read A with a stride of P_A entries across
// the lanes of a warp (one column each): every load instruction touches 18 cache lines for 32 useful values, and those L1
// wavefronts -- 25 x more than the data needs -- bound the kernel (0.70 ms = 0.32 of HBM for the 2048^2 Laplacian squared).
// Here a CTA owns up to 256 consecutive columns j0.. of ONE block column J of C and first copies, fully coalesced, the
// contiguous column ranges it needs -- its own columns of B, and for each of the NSB block columns N = J-uB.. of A the columns
// j0-muB .. j0+nj-1+lamB -- into shared memory; the product then reads shared memory with an odd stride of P entries per lane
// (conflict-free for 8-byte words, one wavefront per instruction for 4-byte ones) and the 256 x P_C result slab is turned
// through shared memory for a coalesced write-back as before.  HBM traffic = every byte of A, B, C once (+ (wB-1)/256 halo).
// Same ascending (N, i) summation order per entry as the other kernels: results are bit-identical.
// one element global -> shared without a register round trip (LDGSTS): all of a thread's copies are in flight at once
__device__ __forceinline__ void cp_async_elem(double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_elem(float* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <typename T, int NSA, int WA, int NSB, int WB>
__global__ void __launch_bounds__(256) bbb_matmul_staged_kernel(const BbbMmParams<T> p, const i64* __restrict__ items, i64 nitems) {
    extern __shared__ __align__(16) unsigned char mmc_smem[];
    constexpr int PA = NSA * WA, PB = NSB * WB, NS = NSA + NSB - 1, NW = WA + WB - 1, TW = 256 + WB - 1;
    T* As = reinterpret_cast<T*>(mmc_smem);            // As[sB][(i - ibase) * PA + rho], ibase = j0 - muB
    T* Bs = As + (size_t)NSB * TW * PA;                // Bs[t * PB + rho]
    T* acc = reinterpret_cast<T*>(mmc_smem);           // acc[rho * 257 + t]: re-uses the staging area after the product
    const int tid = threadIdx.x;
    const int s0 = p.uC - p.uA - p.uB, i0 = p.muC - p.muA - p.muB;   // >= 0 (BandError otherwise)
    for (i64 it = blockIdx.x; it < nitems; it += gridDim.x) {
        const i64 J = items[3 * it], j0 = items[3 * it + 1];
        const int nj = (int)items[3 * it + 2];
        const i64 col0 = p.CB[J] + j0;
        const i64 ibase = j0 - p.muB;
        // 1. stage: B's columns [col0, col0+nj) and, per block column N of A, the columns i in [ibase, ibase + nj + WB - 1)
        // (asynchronous copies: a plain load/store loop kept only ~4 loads in flight per thread and made the staging a
        // chain of memory round trips -- 0.73 ms for the 2048^2 Laplacian squared, no better than the unstaged kernel)
        for (int idx = tid; idx < nj * PB; idx += 256) cp_async_elem(Bs + idx, p.B + (i64)PB * col0 + idx);
#pragma unroll
        for (int sB = 0; sB < NSB; ++sB) {
            const i64 N = J + sB - p.uB;
            if (N < 0 || N >= p.NcA) continue;
            const i64 cN0 = p.CA[N], cN = p.CA[N + 1] - cN0;
            const i64 ilo = max(ibase, (i64)0), ihi = min(ibase + nj + WB - 1, cN);
            if (ihi <= ilo) continue;
            const T* __restrict__ src = p.A + (i64)PA * (cN0 + ilo);
            T* dst = As + (size_t)sB * TW * PA + (size_t)(ilo - ibase) * PA;
            const int cnt = (int)(ihi - ilo) * PA;
            for (int idx = tid; idx < cnt; idx += 256) cp_async_elem(dst + idx, src + idx);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // 2. product: thread t owns column j0 + t (32-bit index arithmetic: block sizes are < 2^30)
        T r[NS][NW];
#pragma unroll
        for (int a = 0; a < NS; ++a)
#pragma unroll
            for (int b = 0; b < NW; ++b) r[a][b] = T(0);
        if (tid < nj) {
            const int j = (int)j0 + tid;
            int rK[NS];   // sizes of the block rows K = J - uA - uB + a of A (0 outside the matrix)
#pragma unroll
            for (int a = 0; a < NS; ++a) {
                const i64 K = J - p.uA - p.uB + a;
                rK[a] = (K >= 0 && K < p.NrA) ? (int)(p.RA[K + 1] - p.RA[K]) : 0;
            }
            const T* bcol = Bs + tid * PB;
#pragma unroll
            for (int sB = 0; sB < NSB; ++sB) {
                const i64 N = J + sB - p.uB;
                if (N < 0 || N >= p.NcA) continue;
                const int cN = (int)(p.CA[N + 1] - p.CA[N]);
#pragma unroll
                for (int iB = 0; iB < WB; ++iB) {
                    const int i = j - p.muB + iB;
                    if ((unsigned)i >= (unsigned)cN) continue;
                    const T b = bcol[sB * WB + iB];
                    const T* acol = As + (size_t)sB * TW * PA + (size_t)(tid + iB) * PA;   // column i = ibase + tid + iB
#pragma unroll
                    for (int sA = 0; sA < NSA; ++sA) {
#pragma unroll
                        for (int iA = 0; iA < WA; ++iA) {
                            const int k = i - p.muA + iA;
                            if ((unsigned)k < (unsigned)rK[sA + sB]) r[sA + sB][iA + iB] = fma(acol[sA * WA + iA], b, r[sA + sB][iA + iB]);
                        }
                    }
                }
            }
        }
        __syncthreads();   // everybody is done with the staged operands: the area becomes the result slab
        for (int rho = 0; rho < p.PC; ++rho) acc[rho * 257 + tid] = T(0);
#pragma unroll
        for (int a = 0; a < NS; ++a)
#pragma unroll
            for (int b = 0; b < NW; ++b) acc[((s0 + a) * p.wC + i0 + b) * 257 + tid] = r[a][b];
        __syncthreads();
        // coalesced write-back of the nj x P_C slab; (t, rho) advance without a division: idx += 256 is t += 256 / P_C,
        // rho += 256 % P_C with one carry
        T* __restrict__ cslab = p.C + (i64)p.PC * col0;
        const int PC = p.PC, dt = 256 / PC, dr = 256 % PC, total = nj * PC;
        int t = tid / PC, rho = tid % PC;
        for (int idx = tid; idx < total; idx += 256) {
            T v = p.alpha * acc[rho * 257 + t];
            if (p.beta != T(0)) v = fma(p.beta, cslab[idx], v);
            cslab[idx] = v;
            t += dt; rho += dr;
            if (rho >= PC) { rho -= PC; ++t; }
        }
        __syncthreads();
    }
}

template <typename T>
int32_t bbb_matmul_impl(bbm_handle_t h, bbm_bbb_desc_t A, bbm_bbb_desc_t B, bbm_bbb_desc_t Cd, T alpha, const T* dA, const T* dB, T beta, T* dC) {
    BBM_REQUIRE_HANDLE(h);
    if (!desc_valid(A) || !desc_valid(B) || !desc_valid(Cd)) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BandedBlockBandedMatrix descriptor");
    if (A->Nc != B->Nr || Cd->Nr != A->Nr || Cd->Nc != B->Nc || A->c != B->r || A->r != Cd->r || B->c != Cd->c)
        return bbm_fail(h, BBM_ERR_DIM, "DimensionMismatch: block axes of A, B and C are incompatible");
    if (A->P > 0 && B->P > 0 && (A->l + B->l > Cd->l || A->u + B->u > Cd->u || A->lam + B->lam > Cd->lam || A->mu + B->mu > Cd->mu))
        return bbm_fail(h, BBM_ERR_BAND, "BandError: the bands of A*B ((%lld,%lld),(%lld,%lld)) do not fit C",
                        (long long)(A->l + B->l), (long long)(A->u + B->u), (long long)(A->lam + B->lam), (long long)(A->mu + B->mu));
    const i64 total = Cd->P * Cd->n;
    if (total == 0) return BBM_OK;
    if (!dC || (A->P * A->n > 0 && !dA) || (B->P * B->n > 0 && !dB)) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    if (dC == dA || dC == dB) return bbm_fail(h, BBM_ERR_ARG, "mul!: destination must not alias an operand");
    BBM_CUDA(h, cudaSetDevice(h->device));
    BbbMmParams<T> p{};
    p.A = dA; p.B = dB; p.C = dC;
    p.RA = A->dR; p.CA = A->dC; p.CB = B->dC;
    p.NrA = A->Nr; p.NcA = A->Nc; p.NcB = B->Nc; p.nC = Cd->n;
    p.lA = (int)A->l; p.uA = (int)A->u; p.lamA = (int)A->lam; p.muA = (int)A->mu; p.wA = (int)std::max<i64>(0, A->lam + A->mu + 1); p.PA = (int)A->P;
    p.lB = (int)B->l; p.uB = (int)B->u; p.lamB = (int)B->lam; p.muB = (int)B->mu; p.wB = (int)std::max<i64>(0, B->lam + B->mu + 1); p.PB = (int)B->P;
    p.lC = (int)Cd->l; p.uC = (int)Cd->u; p.lamC = (int)Cd->lam; p.muC = (int)Cd->mu; p.wC = (int)std::max<i64>(1, Cd->lam + Cd->mu + 1); p.PC = (int)Cd->P;
    p.alpha = alpha; p.beta = beta;
    cudaError_t e = cudaSuccess;
    const size_t smem = (size_t)p.PC * 257 * sizeof(T);
    if (smem <= std::min<size_t>(h->smem_optin, 200 * 1024) && bbm_opt(h, "bbb.mm_kernel", 0) != 2) {
        const i64 blocks = std::min<i64>((Cd->n + 255) / 256, (i64)h->num_sms * 8);
        const bool fixed33 = bbm_opt(h, "bbb.mm_kernel", 0) == 0 && A->P == 9 && B->P == 9 && p.wA == 3 && p.wB == 3;   // (1,1),(1,1) x (1,1),(1,1) and kin
        // staged kernel: operands through shared memory; it wants block columns wide enough to fill a 256-column tile
        const bool staged = fixed33 && bbm_opt(h, "bbb.mm_staged", 1) != 0 && Cd->n >= 64 * Cd->Nc;
        if (staged) {
            if (!Cd->d_mm_items) {
                std::vector<i64> items;
                for (i64 J = 0; J < Cd->Nc; ++J)
                    for (i64 j0 = 0; j0 < Cd->c[J]; j0 += 256) { items.push_back(J); items.push_back(j0); items.push_back(std::min<i64>(256, Cd->c[J] - j0)); }
                Cd->n_mm_items = (i64)items.size() / 3;
                cudaError_t em = cudaMalloc(&Cd->d_mm_items, std::max<size_t>(8, items.size() * sizeof(i64)));
                if (em == cudaSuccess && !items.empty()) em = cudaMemcpy(Cd->d_mm_items, items.data(), items.size() * sizeof(i64), cudaMemcpyHostToDevice);
                if (em != cudaSuccess) { cudaGetLastError(); return bbm_fail(h, BBM_ERR_ALLOC, "column tiles: %s", cudaGetErrorString(em)); }
            }
            const size_t stage = sizeof(T) * ((size_t)3 * 258 * 9 + (size_t)256 * 9);
            const size_t sm = std::max(stage, smem);
            if (sm <= std::min<size_t>(h->smem_optin, 200 * 1024)) {
                e = cudaFuncSetAttribute(bbb_matmul_staged_kernel<T, 3, 3, 3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
                // three CTAs of 74 KB only fit with the largest shared-memory carve-out (the default left room for two)
                if (e == cudaSuccess) e = cudaFuncSetAttribute(bbb_matmul_staged_kernel<T, 3, 3, 3, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
                // exactly one wave of resident CTAs, each looping over its share of the column tiles: a grid of 8 CTAs per
                // SM ran 2.67 waves of the 3 that fit, the last one a third empty ("bbb.mm_ctas_per_sm" overrides)
                int occ = 0;
                if (e == cudaSuccess && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bbb_matmul_staged_kernel<T, 3, 3, 3, 3>, 256, sm) != cudaSuccess) { cudaGetLastError(); occ = 0; }
                const i64 cps = bbm_opt(h, "bbb.mm_ctas_per_sm", 0);
                const i64 per_sm = cps > 0 ? cps : (occ > 0 ? occ : 8);
                const unsigned grid = (unsigned)std::min<i64>(Cd->n_mm_items, (i64)h->num_sms * per_sm);
                if (e == cudaSuccess) bbb_matmul_staged_kernel<T, 3, 3, 3, 3><<<grid, 256, sm, h->stream>>>(p, Cd->d_mm_items, Cd->n_mm_items);
                if (e == cudaSuccess) e = cudaGetLastError();
                if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "bbb_matmul_staged_kernel launch: %s", cudaGetErrorString(e));
                h->launches += 1;
                return BBM_OK;
            }
        }
        if (fixed33) {
            e = cudaFuncSetAttribute(bbb_matmul_col_fixed_kernel<T, 3, 3, 3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) bbb_matmul_col_fixed_kernel<T, 3, 3, 3, 3><<<(unsigned)blocks, 256, smem, h->stream>>>(p);
        } else {
            e = cudaFuncSetAttribute(bbb_matmul_col_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) bbb_matmul_col_kernel<T><<<(unsigned)blocks, 256, smem, h->stream>>>(p);
        }
    } else {
        const i64 blocks = std::min<i64>((total + 255) / 256, (i64)h->num_sms * 32);
        bbb_matmul_kernel<T><<<(unsigned)blocks, 256, 0, h->stream>>>(p);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "bbb_matmul_kernel launch: %s", cudaGetErrorString(e));
    h->launches += 1;
    return BBM_OK;
}

}  // namespace

bool bbm_bbb_desc_view(bbm_bbb_desc_t d, BbbDescView* out) {
    if (!desc_valid(d) || !out) return false;
    *out = BbbDescView{d->h, d->Nr, d->Nc, d->l, d->u, d->lam, d->mu, d->m, d->n, d->P, d->maxr,
                       d->r.data(), d->c.data(), d->dR, d->dC};
    return true;
}

extern "C" {

int32_t bbm_bbb_desc_create(bbm_handle_t h, int64_t Nr, int64_t Nc, const int64_t* r, const int64_t* c, int64_t l,
                            int64_t u, int64_t lam, int64_t mu, bbm_bbb_desc_t* out) {
    BBM_REQUIRE_HANDLE(h);
    if (!out) return bbm_fail(h, BBM_ERR_ARG, "null out pointer");
    *out = nullptr;
    if (Nr < 0 || Nc < 0 || (Nr > 0 && !r) || (Nc > 0 && !c)) return bbm_fail(h, BBM_ERR_ARG, "bad block axes");
    if (Nr > INT32_MAX - 1024 || Nc > INT32_MAX - 1024) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "too many blocks");
    const i64 big = i64(1) << 30;
    if (l < -big || l > big || u < -big || u > big || lam < -big || lam > big || mu < -big || mu > big)
        return bbm_fail(h, BBM_ERR_ARG, "bandwidth out of range");
    bbm_bbb_desc* d = new (std::nothrow) bbm_bbb_desc();
    if (!d) return bbm_fail(h, BBM_ERR_ALLOC, "out of host memory");
    d->h = h; d->Nr = Nr; d->Nc = Nc; d->l = l; d->u = u; d->lam = lam; d->mu = mu;
    d->r.assign(r, r + Nr); d->c.assign(c, c + Nc);
    d->R.assign(Nr + 1, 0); d->C.assign(Nc + 1, 0);
    for (i64 K = 0; K < Nr; ++K) {
        if (r[K] < 0 || r[K] > INT32_MAX / 2) { delete d; return bbm_fail(h, BBM_ERR_ARG, "bad row block size"); }
        d->R[K + 1] = d->R[K] + r[K];
        d->maxr = std::max(d->maxr, r[K]);
    }
    for (i64 J = 0; J < Nc; ++J) {
        if (c[J] < 0 || c[J] > INT32_MAX / 2) { delete d; return bbm_fail(h, BBM_ERR_ARG, "bad column block size"); }
        d->C[J + 1] = d->C[J] + c[J];
    }
    d->m = d->R[Nr]; d->n = d->C[Nc];
    // rows of `data`: max(0,l+u+1)*max(0,lam+mu+1)  (src/BandedBlockBandedMatrix.jl:50)
    const i64 nb = std::max<i64>(0, l + u + 1), w = std::max<i64>(0, lam + mu + 1);
    if (nb > 0 && w > 0 && nb > (i64(1) << 30) / w) { delete d; return bbm_fail(h, BBM_ERR_UNSUPPORTED, "band storage too tall"); }
    d->P = nb * w;
    cudaError_t e = cudaSetDevice(h->device);
    if (e == cudaSuccess) e = cudaMalloc(&d->dR, sizeof(i64) * (size_t)(Nr + 1));
    if (e == cudaSuccess) e = cudaMalloc(&d->dC, sizeof(i64) * (size_t)(Nc + 1));
    if (e == cudaSuccess) e = cudaMemcpy(d->dR, d->R.data(), sizeof(i64) * (size_t)(Nr + 1), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d->dC, d->C.data(), sizeof(i64) * (size_t)(Nc + 1), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (d->dR) cudaFree(d->dR);
        if (d->dC) cudaFree(d->dC);
        delete d;
        return bbm_fail(h, BBM_ERR_CUDA, "descriptor upload: %s", cudaGetErrorString(e));
    }
    *out = d;
    return BBM_OK;
}

int32_t bbm_bbb_desc_destroy(bbm_bbb_desc_t d) {
    if (!d || d->magic != 0xBBBDE5C0u) return BBM_ERR_HANDLE;
    if (bbm_valid(d->h)) cudaSetDevice(d->h->device);
    for (auto& kv : d->parts)
        if (kv.second.d_items) cudaFree(kv.second.d_items);
    for (auto& ch : d->chunks)
        if (ch.sub) bbm_bbb_desc_destroy(ch.sub);
    for (auto& e : d->chunk_ev)
        if (e) cudaEventDestroy(e);
    if (d->dR) cudaFree(d->dR);
    if (d->dC) cudaFree(d->dC);
    if (d->d_mm_items) cudaFree(d->d_mm_items);
    d->magic = 0;
    delete d;
    return BBM_OK;
}

int32_t bbm_bbb_desc_info(bbm_bbb_desc_t d, int64_t* m, int64_t* n, int64_t* P, int64_t* work_items) {
    if (!desc_valid(d)) return BBM_ERR_HANDLE;
    if (m) *m = d->m;
    if (n) *n = d->n;
    if (P) *P = d->P;
    if (work_items) {
        *work_items = 0;
        for (auto& kv : d->parts) *work_items = std::max<i64>(*work_items, kv.second.nitems);
    }
    return BBM_OK;
}

int32_t bbm_bbb_mul_f64(bbm_handle_t h, bbm_bbb_desc_t d, double alpha, const double* d_data, const double* d_X,
                        int64_t ldx, int64_t nrhs, double beta, double* d_Y, int64_t ldy) {
    return bbb_mul_impl<double>(h, d, alpha, d_data, d_X, ldx, nrhs, beta, d_Y, ldy);
}
int32_t bbm_bbb_mul_f32(bbm_handle_t h, bbm_bbb_desc_t d, float alpha, const float* d_data, const float* d_X,
                        int64_t ldx, int64_t nrhs, float beta, float* d_Y, int64_t ldy) {
    return bbb_mul_impl<float>(h, d, alpha, d_data, d_X, ldx, nrhs, beta, d_Y, ldy);
}
int32_t bbm_bbb_trimul_f64(bbm_handle_t h, bbm_bbb_desc_t d, int32_t uplo, int32_t unit, double alpha, const double* d_data,
                           const double* d_X, int64_t ldx, int64_t nrhs, double beta, double* d_Y, int64_t ldy) {
    if (uplo != 'U' && uplo != 'L' && uplo != 'u' && uplo != 'l') return bbm_valid(h) ? bbm_fail(h, BBM_ERR_ARG, "uplo must be 'U' or 'L'") : BBM_ERR_HANDLE;
    return bbb_mul_impl<double>(h, d, alpha, d_data, d_X, ldx, nrhs, beta, d_Y, ldy, (uplo == 'U' || uplo == 'u') ? 1 : 2, unit ? 1 : 0);
}
int32_t bbm_bbb_trimul_f32(bbm_handle_t h, bbm_bbb_desc_t d, int32_t uplo, int32_t unit, float alpha, const float* d_data,
                           const float* d_X, int64_t ldx, int64_t nrhs, float beta, float* d_Y, int64_t ldy) {
    if (uplo != 'U' && uplo != 'L' && uplo != 'u' && uplo != 'l') return bbm_valid(h) ? bbm_fail(h, BBM_ERR_ARG, "uplo must be 'U' or 'L'") : BBM_ERR_HANDLE;
    return bbb_mul_impl<float>(h, d, alpha, d_data, d_X, ldx, nrhs, beta, d_Y, ldy, (uplo == 'U' || uplo == 'u') ? 1 : 2, unit ? 1 : 0);
}
int32_t bbm_bbb_mul_t_f64(bbm_handle_t h, bbm_bbb_desc_t d, double alpha, const double* d_data, const double* d_X, int64_t ldx,
                          int64_t nrhs, double beta, double* d_Y, int64_t ldy) {
    return bbb_mul_t_impl<double>(h, d, alpha, d_data, d_X, ldx, nrhs, beta, d_Y, ldy);
}
int32_t bbm_bbb_mul_t_f32(bbm_handle_t h, bbm_bbb_desc_t d, float alpha, const float* d_data, const float* d_X, int64_t ldx,
                          int64_t nrhs, float beta, float* d_Y, int64_t ldy) {
    return bbb_mul_t_impl<float>(h, d, alpha, d_data, d_X, ldx, nrhs, beta, d_Y, ldy);
}
int32_t bbm_bbb_matmul_f64(bbm_handle_t h, bbm_bbb_desc_t A, bbm_bbb_desc_t B, bbm_bbb_desc_t C, double alpha, const double* d_A,
                           const double* d_B, double beta, double* d_C) {
    return bbb_matmul_impl<double>(h, A, B, C, alpha, d_A, d_B, beta, d_C);
}
int32_t bbm_bbb_matmul_f32(bbm_handle_t h, bbm_bbb_desc_t A, bbm_bbb_desc_t B, bbm_bbb_desc_t C, float alpha, const float* d_A,
                           const float* d_B, float beta, float* d_C) {
    return bbb_matmul_impl<float>(h, A, B, C, alpha, d_A, d_B, beta, d_C);
}
int32_t bbm_bbb_mul_host_f64(bbm_handle_t h, bbm_bbb_desc_t d, double alpha, const double* d_data, const double* h_X,
                             int64_t ldx, int64_t nrhs, double beta, double* h_Y, int64_t ldy) {
    return bbb_mul_host_impl<double>(h, d, alpha, d_data, h_X, ldx, nrhs, beta, h_Y, ldy);
}
int32_t bbm_bbb_mul_host_f32(bbm_handle_t h, bbm_bbb_desc_t d, float alpha, const float* d_data, const float* h_X,
                             int64_t ldx, int64_t nrhs, float beta, float* h_Y, int64_t ldy) {
    return bbb_mul_host_impl<float>(h, d, alpha, d_data, h_X, ldx, nrhs, beta, h_Y, ldy);
}

}  // extern "C"
