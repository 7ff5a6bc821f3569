// BlockSkylineMatrix / BlockBandedMatrix: descriptor, matvec (panel GEMV), block-banded x
// block-banded product (FP64 tensor-core DMMA tiles) and block triangular solve.
//
// Storage (src/BlockSkylineMatrix.jl:6-43, 93-104, 459-479; docs/src/index.md:28-50): for block
// column J the in-band block rows klo[J]..khi[J] form ONE dense column-major panel of
// S_J = R[khi+1]-R[klo] rows x c[J] columns; panels are concatenated in J order.  Block (K,J)
// starts at off[J] + R[K]-R[klo[J]] with leading dimension S_J.  Every stored slot is a matrix
// entry (no unused slots, unlike the banded-block-banded storage).
//
// Reference behaviour being replaced: the [upstream] BlockArrays block loops
// calling one BLAS gemv / gemm / trsv per block on the (ptr, ld) operands above.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <new>

#include "bbm_internal.h"
#include "ptx.cuh"

namespace {

struct SkyEnt {      // one (block or panel piece) operand of a tall-skinny product
    i64 start;       // offset of its first row/column in `data`
    i64 ld;          // leading dimension
    i64 xoff;        // first row of X it multiplies
    i64 ncols;       // columns
};

struct SkyItem {     // one CTA of the panel-GEMV kernel: up to 128 consecutive output rows
    i64 yrow0;       // first output row
    i64 koff;        // row offset added to every entry's start
    int32_t nrows;
    int32_t ent_begin, ent_end;
    int32_t pad;
};

constexpr int kGemvRows = 128;
constexpr int kGemvChunk = 64;

std::atomic<uint64_t> g_serial{1};

}  // namespace

struct bbm_sky_desc {
    uint32_t magic = 0x5C71DE5Cu;
    uint64_t serial = 0;
    bbm_handle_t h = nullptr;
    i64 Nr = 0, Nc = 0, m = 0, n = 0, numel = 0, nblocks = 0;
    std::vector<i64> r, c, R, C, l, u, klo, khi, off, S;
    // matvec tables
    SkyEnt* d_ents = nullptr;
    SkyItem* d_items = nullptr;
    int nitems = 0;
    std::vector<int> row_item_ptr;  // Nr+1: the items of block row K are [row_item_ptr[K], row_item_ptr[K+1])
    // triangular-solve tables, built on first use, per uplo (0 lower, 1 upper)
    struct Tri {
        bool built = false;
        SkyEnt* d_ents = nullptr;      // one panel entry per block column K
        SkyItem* d_items = nullptr;    // update row tiles, grouped by K
        std::vector<int> item_ptr;     // Nr+1
        std::vector<i64> sub_ptr;      // prefix of 32-row sub-blocks per diagonal block
        i64* d_sub = nullptr;          // per sub-block: (data offset of its top-left, ld, size)
        i64* d_place = nullptr;        // per sub-block: (offset in the dense inverse block storage, ld, size)
        i64* d_zero = nullptr;         // rectangles on the zero side of the inverses' triangles: (offset, ld, rows, cols)
        i64 nzero = 0;
        i64 nsub = 0;
        // GEMM-based solve (Float64, many right-hand sides): explicit inverses of the diagonal blocks
        struct Batch { void* cblks = nullptr; void* segs = nullptr; void* tiles = nullptr; int ntiles = 0; bool aligned = false; int tile = 64; };
        bool inv_built = false;
        std::vector<i64> woff, wld;    // inverse of diagonal block K: W + woff[K], leading dimension wld[K]
        i64 wtotal = 0;
        std::vector<Batch> inv_levels; // two batches per level (T = off-diagonal * X22, X12 = -X11 * T)
        i64 step_ldb = -1, step_nrhs = -1, step_ldx = -1;
        Batch stepA, stepB;            // all block steps; tiles of step K are [ptr[K], ptr[K+1])
        std::vector<int> ptrA, ptrB;
        // data-flow solve: ONE persistent kernel walks all block steps; tiles in dependency order with counters
        void* flow_tiles = nullptr;    // FlowTile[flow_ntiles]
        i64 flow_ntiles = 0;
        int* flow_ctr = nullptr;       // 2*Nr completion counters (+ 1 error word), zeroed per call
        i64 flow_ldb = -1, flow_nrhs = -1, flow_ldx = -1, flow_bk = 0, flow_nctr = 0;
        bool flow_ok = false;
        double* d_cond = nullptr;      // per diagonal block: ||T||_inf, ||inv(T)||_inf (condition guard)
    } tri[2];
    // QR-solve plan (Float64): compact-WY form of every panel's reflectors, built on first use
    struct Qr {
        bool built = false;
        i64 ncols = 0;                       // panels = min(Nr, Nc)
        std::vector<i64> srows, vc, vt, g, tt, yt, xx;   // per panel: rows of V; workspace offsets (doubles)
        i64 total = 0, tt_lo = 0, tt_hi = 0; // workspace size; the range holding the T' blocks (zeroed per call)
        i64 maxel = 0;
        i64* d_pan = nullptr;                // per panel: V start in the factors, ld, rows, cols, vc, vt
        i64* d_blk = nullptr;                // per 32-reflector block: G diag offset, ld, T' diag offset, tau offset, size
        i64 nblk = 0;
        Tri::Batch gram, ytb;
        std::vector<Tri::Batch> levels;      // two batches per doubling level of the T factors
        i64 step_ldb = -1, step_nrhs = -1, step_ldw = -1;
        Tri::Batch stepW, stepB;             // the sweep; tiles of panel K are [ptr[K], ptr[K+1])
        std::vector<int> ptrW, ptrB;
        // data-flow sweep: the same two products per panel as tiles of ONE persistent kernel (the solve's sky_trsm_flow_kernel)
        void* flow_tiles = nullptr;
        i64 flow_ntiles = 0, flow_nctr = 0, flow_ldb = -1, flow_nrhs = -1, flow_ldw = -1;
        int* flow_ctr = nullptr;
        bool flow_ok = false;
    } qr;
    // X*A (dense times block-banded): one C block per block column of A; cached for the last (rows of X, ldx, ldy, eltype)
    struct Rmul { i64 p = -1, ldx = -1, ldy = -1; int elem = 0; bool aligned4 = false; Tri::Batch batch; } rmul;
    // descriptor of the row-and-column reversed matrix (its storage is the reversed `data`), for the QL calls
    bbm_sky_desc* rev = nullptr;
    // transposed matvec: one item per group of up to 8 columns of a panel (start, ld = rows, x row start, y row start, columns)
    i64* d_titems = nullptr;
    i64 ntitems = 0;
    // qr! plan (Float64): sub-panels of nb reflectors; tiles of step t are [ptr[t], ptr[t+1])
    struct QrFact {
        bool built = false;
        int nb = 32, variant = 0;
        i64 nsteps = 0, maxrows = 0, ws_total = 0, wlen = 0;
        i64 slotV[2] = {0, 0}, slotY[2] = {0, 0}, slotW[2] = {0, 0};
        i64* d_steps = nullptr;              // per step: slab start, ld, rows, width, tau offset, parity
        Tri::Batch bW, bC;                   // tiles of W = YT C and C -= V W; step t: [ptr[t], mid[t]) on the columns the NEXT
        std::vector<int> ptrW, ptrC;         // step factorises (look-ahead), [mid[t], ptr[t+1]) on everything else
        std::vector<int> midW, midC;
    } qrf;
    // device copy of the structure tables [off(Nc+1) | S(Nc) | klo(Nc) | khi(Nc) | R(Nr+1) | C(Nc+1)] for the
    // element-wise kernels of csrc/structadd.cu, uploaded on first use
    i64* d_struct = nullptr;
    // product plans keyed by the serials of (A, B), owned by the C descriptor
    struct MmPlan;
    std::map<std::pair<uint64_t, uint64_t>, MmPlan*> plans;
};

namespace {

inline bool sky_valid(bbm_sky_desc_t d) { return d && d->magic == 0x5C71DE5Cu && bbm_valid(d->h); }
inline bool sky_has(const bbm_sky_desc* d, i64 K, i64 J) { return K >= d->klo[J] && K <= d->khi[J]; }
inline i64 sky_start(const bbm_sky_desc* d, i64 K, i64 J) { return d->off[J] + d->R[K] - d->R[d->klo[J]]; }

template <typename V>
int32_t upload(bbm_handle_t h, const std::vector<V>& v, V** out) {
    *out = nullptr;
    if (v.empty()) return BBM_OK;
    cudaError_t e = cudaMalloc(out, v.size() * sizeof(V));
    if (e != cudaSuccess) { cudaGetLastError(); return bbm_fail(h, BBM_ERR_ALLOC, "cudaMalloc(%zu): %s", v.size() * sizeof(V), cudaGetErrorString(e)); }
    BBM_CUDA(h, cudaMemcpy(*out, v.data(), v.size() * sizeof(V), cudaMemcpyHostToDevice));
    return BBM_OK;
}

// ------------------------------------------------------------------------------------------------
// panel GEMV:  Y[rows] <- alpha * sum_e A_e[rows, :] * X[xoff_e .. , q] + beta * Y[rows]
// One thread per output row (consecutive rows of a column-major panel are contiguous, so every
// column read is a coalesced 1 KB line per CTA), NR right-hand sides per CTA in registers, the X
// chunk staged through shared memory, 8 independent loads in flight per thread.
// Summation order per entry: ascending global column, like the reference's gemv-per-block sweep.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct GemvParams {
    const T* data;
    const T* X;
    T* Y;
    i64 ldx, ldy;
    const SkyItem* items;
    const SkyEnt* ents;
    T alpha, beta;
    int nrhs;
};

// KSPLIT (gridDim.z > 1): CTA z takes the z-th slice of every entry's columns and accumulates its
// partial product with atomic adds -- only used with beta == 1 (the triangular-solve update), where
// it turns a latency-bound 1536-row panel product into several hundred independent CTAs.
template <typename T, int NR, int BATCH>
__global__ void __launch_bounds__(kGemvRows) sky_gemv_kernel(const GemvParams<T> p) {
    __shared__ T xs[kGemvChunk][NR];
    const SkyItem it = p.items[blockIdx.x];
    const int q0 = blockIdx.y * NR;
    const int nq = min(NR, p.nrhs - q0);
    const int tid = threadIdx.x;
    const bool active = tid < it.nrows;
    const int nz = gridDim.z, z = blockIdx.z;
    T acc[NR];
#pragma unroll
    for (int q = 0; q < NR; ++q) acc[q] = T(0);

    for (int e = it.ent_begin; e < it.ent_end; ++e) {
        const SkyEnt en = p.ents[e];
        const T* __restrict__ a = p.data + en.start + it.koff + tid;
        // column slice of this CTA (multiples of the chunk so that slices stay aligned)
        i64 jbeg = 0, jend = en.ncols;
        if (nz > 1) {
            const i64 per = ((en.ncols + nz - 1) / nz + kGemvChunk - 1) / kGemvChunk * kGemvChunk;
            jbeg = min(en.ncols, per * z);
            jend = min(en.ncols, jbeg + per);
        }
        for (i64 j0 = jbeg; j0 < jend; j0 += kGemvChunk) {
            const int nj = (int)min((i64)kGemvChunk, jend - j0);
            __syncthreads();
            for (int idx = tid; idx < NR * kGemvChunk; idx += kGemvRows) {
                const int q = idx / kGemvChunk, jj = idx % kGemvChunk;
                xs[jj][q] = (q < nq && jj < nj) ? p.X[(i64)(q0 + q) * p.ldx + en.xoff + j0 + jj] : T(0);
            }
            __syncthreads();
            if (active) {
                const T* __restrict__ aj = a + j0 * en.ld;
                int jj = 0;
                for (; jj + BATCH <= nj; jj += BATCH) {
                    T v[BATCH];
#pragma unroll
                    for (int t = 0; t < BATCH; ++t) v[t] = __ldg(aj + (i64)(jj + t) * en.ld);
#pragma unroll
                    for (int t = 0; t < BATCH; ++t)
#pragma unroll
                        for (int q = 0; q < NR; ++q) acc[q] = fma(v[t], xs[jj + t][q], acc[q]);
                }
                for (; jj < nj; ++jj) {
                    const T v = __ldg(aj + (i64)jj * en.ld);
#pragma unroll
                    for (int q = 0; q < NR; ++q) acc[q] = fma(v, xs[jj][q], acc[q]);
                }
            }
        }
    }
    if (active) {
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            if (q < nq) {
                T* yp = p.Y + (i64)(q0 + q) * p.ldy + it.yrow0 + tid;
                if (nz > 1) {
                    atomicAdd(yp, p.alpha * acc[q]);
                } else {
                    T y = p.alpha * acc[q];
                    if (p.beta != T(0)) y = fma(p.beta, *yp, y);
                    *yp = y;
                }
            }
        }
    }
}

template <typename T>
int32_t launch_gemv(bbm_handle_t h, const GemvParams<T>& p, int nitems, int nrhs, int ksplit = 1) {
    if (nitems <= 0 || nrhs <= 0) return BBM_OK;
    if (ksplit > 1 && p.beta != T(1)) return bbm_fail(h, BBM_ERR_ARG, "internal: split products need beta == 1");
    cudaError_t e;
    const unsigned z = (unsigned)std::max(1, ksplit);
    if (nrhs == 1) {
        sky_gemv_kernel<T, 1, 8><<<dim3(nitems, 1, z), kGemvRows, 0, h->stream>>>(p);
    } else if (nrhs <= 4) {
        sky_gemv_kernel<T, 4, 8><<<dim3(nitems, 1, z), kGemvRows, 0, h->stream>>>(p);
    } else if (nrhs <= 8 || ksplit <= 1) {
        sky_gemv_kernel<T, 8, 8><<<dim3(nitems, (nrhs + 7) / 8, z), kGemvRows, 0, h->stream>>>(p);
    } else {
        sky_gemv_kernel<T, 16, 16><<<dim3(nitems, (nrhs + 15) / 16, z), kGemvRows, 0, h->stream>>>(p);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "sky_gemv_kernel launch: %s", cudaGetErrorString(e));
    h->launches += 1;
    return BBM_OK;
}

template <typename T>
__global__ void __launch_bounds__(256) sky_prescale_kernel(T* __restrict__ Y, i64 ldy, i64 m, i64 nrhs, T beta) {
    for (i64 q = blockIdx.y; q < nrhs; q += gridDim.y)
        for (i64 i = (i64)blockIdx.x * 256 + threadIdx.x; i < m; i += (i64)gridDim.x * 256) {
            T* y = Y + i + q * ldy;
            *y = beta == T(0) ? T(0) : beta * *y;   // beta == 0 never reads the destination
        }
}

// Klo/Khi: only the block rows [Klo,Khi) are multiplied (Khi < 0: all).  phase: 0 = one self-contained product;
// 1 = first part of a product issued in pieces (does the beta pre-scale of ALL rows if the split kernel is used);
// 2 = a later piece (no pre-scale; split decision taken from the whole matrix so that the pieces agree).
template <typename T>
int32_t sky_mul_impl(bbm_handle_t h, bbm_sky_desc_t d, T alpha, const T* d_data, const T* d_X, i64 ldx, i64 nrhs, T beta,
                     T* d_Y, i64 ldy, i64 Klo = 0, i64 Khi = -1, int phase = 0) {
    BBM_REQUIRE_HANDLE(h);
    if (!sky_valid(d)) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BlockSkylineMatrix descriptor");
    if (nrhs < 0) return bbm_fail(h, BBM_ERR_ARG, "nrhs < 0");
    if (nrhs > 0 && (ldx < std::max<i64>(1, d->n) || ldy < std::max<i64>(1, d->m)))
        return bbm_fail(h, BBM_ERR_DIM, "DimensionMismatch: A is %lld x %lld, ldx=%lld, ldy=%lld", (long long)d->m,
                        (long long)d->n, (long long)ldx, (long long)ldy);
    if (d->m == 0 || nrhs == 0) return BBM_OK;
    if (!d_Y || (d->n > 0 && !d_X) || (d->numel > 0 && !d_data)) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    if (nrhs > 8 * 65535) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "too many right-hand sides");
    BBM_CUDA(h, cudaSetDevice(h->device));
    if (Khi < 0) Khi = d->Nr;
    if (Klo < 0 || Khi > d->Nr || Klo > Khi) return bbm_fail(h, BBM_ERR_ARG, "bad block-row range");
    const int it0 = d->row_item_ptr[Klo], it1 = d->row_item_ptr[Khi];
    GemvParams<T> p{d_data, d_X, d_Y, ldx, ldy, d->d_items + it0, d->d_ents, alpha, beta, (int)nrhs};
    // Few row tiles (a small matrix, or one rank's share of a large one) leave most of an SM's warp slots empty and
    // the stream latency-bound: split every panel's columns over gridDim.z CTAs that accumulate with atomics into a
    // pre-scaled y (the same split the triangular solve's update uses).
    const i64 ycols = nrhs == 1 ? 1 : (nrhs <= 4 ? 1 : (nrhs + 7) / 8);
    // measured (256-row blocks, (2,2), one vector): a CTA's column loop alone takes ~190 us whatever the matrix size,
    // so up to ~1000 row tiles a 4-way split is 1.2x - 3x faster (164 MB: 191 -> 60 us, 668 MB: 206 -> 110 us,
    // 1.3 GB: 239 -> 203 us); at 4096 tiles (cfg4's operand, 7.07 TB/s) the unsplit kernel already saturates HBM
    i64 z = bbm_opt(h, "sky.mv_split", 0);
    if (z <= 0) z = (i64)d->nitems * ycols <= 14 * (i64)h->num_sms ? 4 : ((i64)d->nitems * ycols <= 21 * (i64)h->num_sms ? 2 : 1);
    if (z >= 2 && d->nitems > 0) {
        if (phase != 2) {
            const unsigned gx = (unsigned)std::min<i64>(65535, (d->m + 255) / 256);
            sky_prescale_kernel<T><<<dim3(gx, (unsigned)std::min<i64>(nrhs, 65535)), 256, 0, h->stream>>>(d_Y, ldy, d->m, nrhs, beta);
            h->launches += 1;
        }
        p.beta = T(1);
        return launch_gemv<T>(h, p, it1 - it0, (int)nrhs, (int)z);
    }
    return launch_gemv<T>(h, p, it1 - it0, (int)nrhs);
}

// ------------------------------------------------------------------------------------------------
// block-banded x block-banded product:  C <- alpha*A*B + beta*C
// Work item = one BM x BN tile of one stored block (K,J) of C; its inner dimension is the
// concatenation of the segments N in colsupport(B,J) with K in colsupport(A,N)
// ([upstream] _block_muladd! triple loop; result structure src/linalg.jl:27-48).  Every tile of C
// is produced by exactly one CTA and written once: no atomics, no beta pre-pass.
// FP64 runs on the tensor cores: mma.sync.m8n8k4.f64 (SASS DMMA) fed from a 3-stage cp.async
// shared-memory ring -- tcgen05 has no f64 kind, so this is the sm_100a tensor path for doubles.
// ------------------------------------------------------------------------------------------------
struct MmCBlk {
    i64 cstart, ldc;
    int32_t rK, cJ;
    int32_t seg_begin, seg_end;
};
struct MmSeg {
    i64 astart, lda, bstart, ldb, kc;
};
struct MmTile {
    int32_t cblk, ti, tj, pad;
};

template <typename T>
struct MmParams {
    const T* A;
    const T* B;
    T* C;
    const MmCBlk* cblks;
    const MmSeg* segs;
    const MmTile* tiles;
    T alpha, beta;
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src, int bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int BM, int BN, int kMmBK, int kMmStages = 3>
constexpr size_t mm_smem_bytes() {
    return (size_t)kMmStages * ((size_t)kMmBK * (BM + 4) + (size_t)BN * (kMmBK + 4)) * sizeof(double);
}

template <int BM, int BN, int WM, int WN, int kMmBK, bool ALIGNED, bool ATOMIC = false, int kMmStages = 3>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32) sky_dgemm_dmma_kernel(const MmParams<double> p) {
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    constexpr int LDA = BM + 4;     // As[k][m]: 4 k-rows of a fragment land in distinct bank groups
    constexpr int LDB = kMmBK + 4;  // Bs[n][k]
    constexpr int MF = WM / 8, NF = WN / 8;
    extern __shared__ __align__(16) unsigned char mm_smem[];
    double* As = reinterpret_cast<double*>(mm_smem);
    double* Bs = As + (size_t)kMmStages * kMmBK * LDA;

    // the tile tables are written once at plan time: they may be read before the predecessor kernel (a
    // programmatic dependent launch of the triangular solve's chain) has finished; operands may not
    const MmTile tile = p.tiles[blockIdx.x];
    const MmCBlk cb = p.cblks[tile.cblk];
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm0 = (warp % (BM / WM)) * WM, wn0 = (warp / (BM / WM)) * WN;
    const int row0 = tile.ti * BM, col0 = tile.tj * BN;

    // flattened k-chunk iterator of the producer
    int pseg = cb.seg_begin;
    i64 pk0 = 0;
    int total = 0;
    for (int s = cb.seg_begin; s < cb.seg_end; ++s) total += (int)((p.segs[s].kc + kMmBK - 1) / kMmBK);

    auto load_chunk = [&](int stage) {
        const MmSeg sg = p.segs[pseg];
        double* as = As + (size_t)stage * kMmBK * LDA;
        double* bs = Bs + (size_t)stage * BN * LDB;
        if (ALIGNED) {
            for (int v = tid; v < (BM / 2) * kMmBK; v += NT) {
                const int kk = v / (BM / 2), i = 2 * (v % (BM / 2));
                const int row = row0 + i;
                int nv = cb.rK - row; nv = nv < 0 ? 0 : (nv > 2 ? 2 : nv);
                const bool kv = pk0 + kk < sg.kc;
                const int bytes = kv ? 8 * nv : 0;
                const double* src = bytes ? p.A + sg.astart + row + (pk0 + kk) * sg.lda : p.A;
                cp_async16(as + kk * LDA + i, src, bytes);
            }
            for (int v = tid; v < BN * (kMmBK / 2); v += NT) {
                const int j = v / (kMmBK / 2), kk = 2 * (v % (kMmBK / 2));
                const int col = col0 + j;
                i64 nv = sg.kc - (pk0 + kk); nv = nv < 0 ? 0 : (nv > 2 ? 2 : nv);
                const int bytes = col < cb.cJ ? 8 * (int)nv : 0;
                const double* src = bytes ? p.B + sg.bstart + (pk0 + kk) + (i64)col * sg.ldb : p.B;
                cp_async16(bs + j * LDB + kk, src, bytes);
            }
        } else {
            for (int v = tid; v < BM * kMmBK; v += NT) {
                const int kk = v / BM, i = v % BM;
                const int row = row0 + i;
                const int bytes = (row < cb.rK && pk0 + kk < sg.kc) ? 8 : 0;
                const double* src = bytes ? p.A + sg.astart + row + (pk0 + kk) * sg.lda : p.A;
                cp_async8(as + kk * LDA + i, src, bytes);
            }
            for (int v = tid; v < BN * kMmBK; v += NT) {
                const int j = v / kMmBK, kk = v % kMmBK;
                const int col = col0 + j;
                const int bytes = (col < cb.cJ && pk0 + kk < sg.kc) ? 8 : 0;
                const double* src = bytes ? p.B + sg.bstart + (pk0 + kk) + (i64)col * sg.ldb : p.B;
                cp_async8(bs + j * LDB + kk, src, bytes);
            }
        }
        pk0 += kMmBK;
        if (pk0 >= sg.kc) { pk0 = 0; ++pseg; }
    };

    double acc[MF][NF][2];
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    int issued = 0;
    for (; issued < kMmStages - 1; ++issued) {
        if (issued < total) load_chunk(issued % kMmStages);
        cp_async_commit();
    }
    const int ar = lane >> 2, ac = lane & 3;  // fragment coordinates of this lane
    for (int c = 0; c < total; ++c) {
        cp_async_wait<kMmStages - 2>();
        __syncthreads();
        const double* as = As + (size_t)(c % kMmStages) * kMmBK * LDA;
        const double* bs = Bs + (size_t)(c % kMmStages) * BN * LDB;
#pragma unroll
        for (int k4 = 0; k4 < kMmBK; k4 += 4) {
            // the next chunk's cp.async issue (address arithmetic, ~100 instructions) is staggered: the
            // first half of the warps issues after the 1st k-step, the second half mid-chunk, so that on
            // every scheduler one warp keeps the DMMA pipe fed while the other issues loads
            if ((k4 == 4 && warp < NT / 64) || (k4 == 4 * (kMmBK / 8) + 4 && warp >= NT / 64)) {
                if (issued < total) load_chunk(issued % kMmStages);
                cp_async_commit();
                ++issued;
            }
            double af[MF], bf[NF];
#pragma unroll
            for (int i = 0; i < MF; ++i) af[i] = as[(k4 + ac) * LDA + wm0 + i * 8 + ar];
#pragma unroll
            for (int j = 0; j < NF; ++j) bf[j] = bs[(wn0 + j * 8 + ar) * LDB + k4 + ac];
#pragma unroll
            for (int i = 0; i < MF; ++i)
#pragma unroll
                for (int j = 0; j < NF; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the successor may start its prologue

    // epilogue: accumulator (row = lane/4, cols = 2*(lane%4)+{0,1}) -> C, written once
    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int i = 0; i < MF; ++i) {
        const int row = row0 + wm0 + i * 8 + ar;
        if (row >= cb.rK) continue;
#pragma unroll
        for (int j = 0; j < NF; ++j) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int col = col0 + wn0 + j * 8 + 2 * ac + t;
                if (col < cb.cJ) {
                    double* cp = p.C + cb.cstart + row + (i64)col * cb.ldc;
                    double v = alpha * acc[i][j][t];
                    if (ATOMIC) {
                        atomicAdd(cp, v);  // split-k partial product (beta == 1 by construction)
                    } else {
                        if (beta != 0.0) v = fma(beta, *cp, v);
                        *cp = v;
                    }
                }
            }
        }
    }
}


// Warp-specialised variant for full tiles (every block a multiple of 128 x 128 with inner dimensions a multiple of 32, 16-byte
// aligned operands -- cfg4): eight DMMA warps consume a 3-stage ring that one producer WARP fills with 1-D bulk copies
// (cp.async.bulk, SASS UBLKCP: 32 copies of 1 KB for the A chunk -- a k-column of the tile is contiguous in the column-major
// panel -- and 128 copies of 256 B for the B chunk, five per lane), completion counted in bytes on the stage's `full` mbarrier; a warp hands a
// stage back with one arrive on its `empty` mbarrier.  No thread of the DMMA warps computes a global address, and there is no
// CTA-wide barrier inside the k loop: warps drift apart by up to a stage instead of meeting twice per chunk.  2-D tensor maps
// (cp.async.bulk.tensor) would need one descriptor per panel (every panel has its own leading dimension); the padded shared-memory
// rows the fragments want (LDA = 132, LDB = 36 doubles: bank-conflict-free) are 16-byte multiples, which is all the 1-D copies need.
template <int BM, int BN, int BK, int STAGES>
__global__ void __launch_bounds__(288, 1) sky_dgemm_dmma_ws_kernel(const MmParams<double> p) {
    constexpr int WM = 64, WN = 32, MF = WM / 8, NF = WN / 8;
    constexpr int LDA = BM + 4, LDB = BK + 4;
    extern __shared__ __align__(128) unsigned char mm_smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(mm_smem);            // [STAGES]
    uint64_t* empty = full + STAGES;                                  // [STAGES]
    double* As = reinterpret_cast<double*>(mm_smem + 128);
    double* Bs = As + (size_t)STAGES * BK * LDA;
    const MmTile tile = p.tiles[blockIdx.x];
    const MmCBlk cb = p.cblks[tile.cblk];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 8); }
        ptx::fence_mbar_init();
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int row0 = tile.ti * BM, col0 = tile.tj * BN;
    int total = 0;
    for (int s = cb.seg_begin; s < cb.seg_end; ++s) total += (int)(p.segs[s].kc / BK);

    if (warp == 8) {
        // ---- producer warp: lane 0 arms the stage's barrier, then all 32 lanes issue their share of the 160 bulk copies
        // (a single issuing thread was the bottleneck of the first version: 79 ms against 54 ms for the cp.async kernel) ----
        int seg = cb.seg_begin;
        i64 k0 = 0;
        constexpr uint32_t kBytes = (uint32_t)((BM * BK + BN * BK) * sizeof(double));
        for (int c = 0; c < total; ++c) {
            const int st = c % STAGES;
            if (c >= STAGES) ptx::mbar_wait(&empty[st], (uint32_t)(((c / STAGES) - 1) & 1));
            const MmSeg sg = p.segs[seg];
            double* as = As + (size_t)st * BK * LDA;
            double* bs = Bs + (size_t)st * BN * LDB;
            if (lane == 0) ptx::mbar_arrive_expect_tx(&full[st], kBytes);
            __syncwarp();
            const double* ag = p.A + sg.astart + row0 + k0 * sg.lda;
            for (int kk = lane; kk < BK; kk += 32) ptx::bulk_g2s(as + kk * LDA, ag + (i64)kk * sg.lda, BM * sizeof(double), &full[st]);
            const double* bg = p.B + sg.bstart + k0 + (i64)col0 * sg.ldb;
#pragma unroll
            for (int j = lane; j < BN; j += 32) ptx::bulk_g2s(bs + j * LDB, bg + (i64)j * sg.ldb, BK * sizeof(double), &full[st]);
            k0 += BK;
            if (k0 >= sg.kc) { k0 = 0; ++seg; }
        }
        return;
    }
    // ---- consumers: 8 warps (2 x 4), warp tile 64 x 32 ----
    const int wm0 = (warp % (BM / WM)) * WM, wn0 = (warp / (BM / WM)) * WN;
    const int ar = lane >> 2, ac = lane & 3;
    double acc[MF][NF][2];
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int c = 0; c < total; ++c) {
        const int st = c % STAGES;
        ptx::mbar_wait(&full[st], (uint32_t)((c / STAGES) & 1));
        const double* as = As + (size_t)st * BK * LDA;
        const double* bs = Bs + (size_t)st * BN * LDB;
#pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
            double af[MF], bf[NF];
#pragma unroll
            for (int i = 0; i < MF; ++i) af[i] = as[(k4 + ac) * LDA + wm0 + i * 8 + ar];
#pragma unroll
            for (int j = 0; j < NF; ++j) bf[j] = bs[(wn0 + j * 8 + ar) * LDB + k4 + ac];
#pragma unroll
            for (int i = 0; i < MF; ++i)
#pragma unroll
                for (int j = 0; j < NF; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[st]);   // this warp is done with the stage
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int i = 0; i < MF; ++i) {
        const int row = row0 + wm0 + i * 8 + ar;
#pragma unroll
        for (int j = 0; j < NF; ++j) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int col = col0 + wn0 + j * 8 + 2 * ac + t;
                double* cp = p.C + cb.cstart + row + (i64)col * cb.ldc;
                double v = alpha * acc[i][j][t];
                if (beta != 0.0) v = fma(beta, *cp, v);
                *cp = v;
            }
        }
    }
}

template <int BM, int BN, int BK, int STAGES>
constexpr size_t mm_ws_smem_bytes() {
    return 128 + (size_t)STAGES * ((size_t)BK * (BM + 4) + (size_t)BN * (BK + 4)) * sizeof(double);
}

// SIMT tile kernel: Float32 (tcgen05 kind::tf32 would drop 13 mantissa bits, outside the
// 64*eps tolerance) and a cross-check path for Float64 ("sky.mm_kernel" = 2).
template <typename T>
__global__ void __launch_bounds__(256) sky_gemm_simt_kernel(const MmParams<T> p) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ T As[BK][BM + 1];
    __shared__ T Bs[BK][BN + 1];
    const MmTile tile = p.tiles[blockIdx.x];
    const MmCBlk cb = p.cblks[tile.cblk];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int row0 = tile.ti * BM, col0 = tile.tj * BN;
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
    for (int s = cb.seg_begin; s < cb.seg_end; ++s) {
        const MmSeg sg = p.segs[s];
        for (i64 k0 = 0; k0 < sg.kc; k0 += BK) {
            __syncthreads();
            for (int v = tid; v < BM * BK; v += 256) {
                const int kk = v / BM, i = v % BM;
                const int row = row0 + i;
                As[kk][i] = (row < cb.rK && k0 + kk < sg.kc) ? p.A[sg.astart + row + (k0 + kk) * sg.lda] : T(0);
            }
            for (int v = tid; v < BN * BK; v += 256) {
                const int j = v / BK, kk = v % BK;
                const int col = col0 + j;
                Bs[kk][j] = (col < cb.cJ && k0 + kk < sg.kc) ? p.B[sg.bstart + (k0 + kk) + (i64)col * sg.ldb] : T(0);
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                T a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = As[kk][tx + 16 * i];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = Bs[kk][ty + 16 * j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = row0 + tx + 16 * i;
        if (row >= cb.rK) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = col0 + ty + 16 * j;
            if (col < cb.cJ) {
                T* cp = p.C + cb.cstart + row + (i64)col * cb.ldc;
                T v = p.alpha * acc[i][j];
                if (p.beta != T(0)) v = fma(p.beta, *cp, v);
                *cp = v;
            }
        }
    }
}

// Float32 product kernel.  Single-precision products that have to agree with the reference's sgemm need the FP32 FMA pipe
// (tcgen05 kind::tf32 drops 13 mantissa bits; a 3 x TF32 split triples the tensor work and still differs in the last
// bits), so this is a register-tiled SIMT GEMM: 128 x 128 tile, 256 threads x (8 x 8) accumulators, operands through a
// 3-stage cp.async ring in the layouts global memory already has (A: rows contiguous, B: k contiguous), both read from
// shared memory as float4 -- 16 LDS.128 per 256 FFMA.  The first version (64 x 64 tile, 4 x 4 per thread, synchronous
// scalar loads: `sky_gemm_simt_kernel` above) ran cfg4's structure at 24.0 TFLOP/s = 0.375 of cuBLAS SGEMM.
constexpr int kSgBM = 128, kSgBN = 128, kSgBK = 16, kSgStages = 3, kSgLDA = kSgBM + 4, kSgLDB = kSgBK + 4;
constexpr size_t kSgSmem = (size_t)kSgStages * ((size_t)kSgBK * kSgLDA + (size_t)kSgBN * kSgLDB) * sizeof(float);

__device__ __forceinline__ void cp_async4(void* dst, const void* src, int bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes) : "memory");
}

template <bool ALIGNED>
__global__ void __launch_bounds__(256, 2) sky_sgemm_simt_kernel(const MmParams<float> p) {
    extern __shared__ __align__(16) unsigned char mm_smem[];
    float* As = reinterpret_cast<float*>(mm_smem);                        // [stage][k][LDA]  (row index contiguous)
    float* Bs = As + (size_t)kSgStages * kSgBK * kSgLDA;                  // [stage][n][LDB]  (k index contiguous)
    const MmTile tile = p.tiles[blockIdx.x];
    const MmCBlk cb = p.cblks[tile.cblk];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = tile.ti * kSgBM, col0 = tile.tj * kSgBN;
    int pseg = cb.seg_begin;
    i64 pk0 = 0;
    int total = 0;
    for (int s = cb.seg_begin; s < cb.seg_end; ++s) total += (int)((p.segs[s].kc + kSgBK - 1) / kSgBK);

    auto load_chunk = [&](int stage) {
        const MmSeg sg = p.segs[pseg];
        float* as = As + (size_t)stage * kSgBK * kSgLDA;
        float* bs = Bs + (size_t)stage * kSgBN * kSgLDB;
        if (ALIGNED) {
            for (int v = tid; v < (kSgBM / 4) * kSgBK; v += 256) {          // A: 4 consecutive rows of one k
                const int kk = v / (kSgBM / 4), i = 4 * (v % (kSgBM / 4));
                const int row = row0 + i;
                int nv = cb.rK - row; nv = nv < 0 ? 0 : (nv > 4 ? 4 : nv);
                const int bytes = pk0 + kk < sg.kc ? 4 * nv : 0;
                const float* src = bytes ? p.A + sg.astart + row + (pk0 + kk) * sg.lda : p.A;
                cp_async16(as + kk * kSgLDA + i, src, bytes);
            }
            for (int v = tid; v < kSgBN * (kSgBK / 4); v += 256) {          // B: 4 consecutive k of one column
                const int j = v / (kSgBK / 4), kk = 4 * (v % (kSgBK / 4));
                const int col = col0 + j;
                i64 nv = sg.kc - (pk0 + kk); nv = nv < 0 ? 0 : (nv > 4 ? 4 : nv);
                const int bytes = col < cb.cJ ? 4 * (int)nv : 0;
                const float* src = bytes ? p.B + sg.bstart + (pk0 + kk) + (i64)col * sg.ldb : p.B;
                cp_async16(bs + j * kSgLDB + kk, src, bytes);
            }
        } else {
            for (int v = tid; v < kSgBM * kSgBK; v += 256) {
                const int kk = v / kSgBM, i = v % kSgBM;
                const int row = row0 + i;
                const int bytes = (row < cb.rK && pk0 + kk < sg.kc) ? 4 : 0;
                const float* src = bytes ? p.A + sg.astart + row + (pk0 + kk) * sg.lda : p.A;
                cp_async4(as + kk * kSgLDA + i, src, bytes);
            }
            for (int v = tid; v < kSgBN * kSgBK; v += 256) {
                const int j = v / kSgBK, kk = v % kSgBK;
                const int col = col0 + j;
                const int bytes = (col < cb.cJ && pk0 + kk < sg.kc) ? 4 : 0;
                const float* src = bytes ? p.B + sg.bstart + (pk0 + kk) + (i64)col * sg.ldb : p.B;
                cp_async4(bs + j * kSgLDB + kk, src, bytes);
            }
        }
        pk0 += kSgBK;
        if (pk0 >= sg.kc) { pk0 = 0; ++pseg; }
    };

    // thread (tx, ty): rows {4 tx .. 4 tx + 3} and {64 + 4 tx ..}, columns {4 ty ..} and {64 + 4 ty ..}
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    int issued = 0;
    for (; issued < kSgStages - 1; ++issued) {
        if (issued < total) load_chunk(issued % kSgStages);
        cp_async_commit();
    }
    for (int c = 0; c < total; ++c) {
        cp_async_wait<kSgStages - 2>();
        __syncthreads();
        if (issued < total) load_chunk(issued % kSgStages);    // refills the stage everybody left before this barrier
        cp_async_commit();
        ++issued;
        const float* as = As + (size_t)(c % kSgStages) * kSgBK * kSgLDA;
        const float* bs = Bs + (size_t)(c % kSgStages) * kSgBN * kSgLDB;
#pragma unroll
        for (int k4 = 0; k4 < kSgBK; k4 += 4) {
            float4 b4[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) b4[j] = *reinterpret_cast<const float4*>(bs + ((j < 4 ? 0 : 64) + 4 * ty + (j & 3)) * kSgLDB + k4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(as + (k4 + kk) * kSgLDA + 4 * tx);
                const float4 a1 = *reinterpret_cast<const float4*>(as + (k4 + kk) * kSgLDA + 64 + 4 * tx);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float b = kk == 0 ? b4[j].x : kk == 1 ? b4[j].y : kk == 2 ? b4[j].z : b4[j].w;
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i][j] = fmaf(a[i], b, acc[i][j]);
                }
            }
        }
    }
    cp_async_wait<0>();
    const float alpha = p.alpha, beta = p.beta;
    const bool vec = ((cb.cstart | cb.ldc) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = col0 + (j < 4 ? 0 : 64) + 4 * ty + (j & 3);
        if (col >= cb.cJ) continue;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const int row = row0 + 64 * h2 + 4 * tx;
            float* cp = p.C + cb.cstart + row + (i64)col * cb.ldc;
            if (vec && row + 3 < cb.rK) {
                float4 v = make_float4(alpha * acc[4 * h2][j], alpha * acc[4 * h2 + 1][j], alpha * acc[4 * h2 + 2][j], alpha * acc[4 * h2 + 3][j]);
                if (beta != 0.0f) {
                    const float4 o = *reinterpret_cast<const float4*>(cp);
                    v.x = fmaf(beta, o.x, v.x); v.y = fmaf(beta, o.y, v.y); v.z = fmaf(beta, o.z, v.z); v.w = fmaf(beta, o.w, v.w);
                }
                *reinterpret_cast<float4*>(cp) = v;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (row + i < cb.rK) {
                        float v = alpha * acc[4 * h2 + i][j];
                        if (beta != 0.0f) v = fmaf(beta, cp[i], v);
                        cp[i] = v;
                    }
                }
            }
        }
    }
}

}  // namespace

struct bbm_sky_desc::MmPlan {
    MmCBlk* d_cblks = nullptr;
    MmSeg* d_segs = nullptr;
    MmTile* d_tiles[2] = {nullptr, nullptr};  // [0] 128x128 DMMA tiles, [1] 64x64 tiles
    int ntiles[2] = {0, 0};
    int ninterior[2] = {0, 0};                // leading tiles whose product segments only use owned block rows of B
    bool aligned = false;
    bool aligned4 = false;                    // all operand offsets and strides multiples of 4 elements (16 bytes of Float32)
    bool full128 = false;                     // every C block a multiple of 128 x 128, inner dimensions multiples of 32, 16-byte strides
    double flops = 0;
};

namespace {

void free_plan(bbm_sky_desc::MmPlan* pl) {
    if (!pl) return;
    if (pl->d_cblks) cudaFree(pl->d_cblks);
    if (pl->d_segs) cudaFree(pl->d_segs);
    for (int i = 0; i < 2; ++i)
        if (pl->d_tiles[i]) cudaFree(pl->d_tiles[i]);
    delete pl;
}

// structure checks + plan construction for C <- A*B
// own_lo/own_hi: the block rows [own_lo, own_hi) of B this rank owns (sharded product: the others are halo rows that
// arrive over NVLink); tiles that only multiply owned rows are ordered first so that they can run while the halo is
// still in flight.  Default: everything is owned.
int32_t build_mm_plan(bbm_handle_t h, bbm_sky_desc_t A, bbm_sky_desc_t B, bbm_sky_desc_t Cd, bbm_sky_desc::MmPlan** out,
                      i64 own_lo = 0, i64 own_hi = INT64_MAX) {
    // block axes must be compatible ([upstream] DimensionMismatch)
    if (A->Nc != B->Nr || Cd->Nr != A->Nr || Cd->Nc != B->Nc || A->c != B->r || A->r != Cd->r || B->c != Cd->c)
        return bbm_fail(h, BBM_ERR_DIM, "DimensionMismatch: block axes of A (%lldx%lld blocks), B (%lldx%lld) and C (%lldx%lld) are incompatible",
                        (long long)A->Nr, (long long)A->Nc, (long long)B->Nr, (long long)B->Nc, (long long)Cd->Nr, (long long)Cd->Nc);
    // every product block must be stored in C (BandError, src/BlockSkylineMatrix.jl:491)
    for (i64 J = 0; J < B->Nc; ++J)
        for (i64 N = B->klo[J]; N <= B->khi[J]; ++N)
            for (i64 K = A->klo[N]; K <= A->khi[N]; ++K)
                if (!sky_has(Cd, K, J) && A->r[K] > 0 && B->c[J] > 0 && A->c[N] > 0)
                    return bbm_fail(h, BBM_ERR_BAND, "BandError: product block (%lld,%lld) lies outside the block band of C", (long long)K, (long long)J);
    std::vector<MmCBlk> cblks;
    std::vector<MmSeg> segs;
    std::vector<MmTile> tiles[2], btiles[2];
    const int tsz[2] = {128, 64};
    bool aligned = true, aligned4 = true, full128 = true;
    double flops = 0;
    for (i64 J = 0; J < Cd->Nc; ++J) {
        if (Cd->c[J] == 0) continue;
        for (i64 K = Cd->klo[J]; K <= Cd->khi[J]; ++K) {
            if (Cd->r[K] == 0) continue;
            MmCBlk cb;
            cb.cstart = sky_start(Cd, K, J); cb.ldc = Cd->S[J];
            cb.rK = (int32_t)Cd->r[K]; cb.cJ = (int32_t)Cd->c[J];
            cb.seg_begin = (int32_t)segs.size();
            bool interior = true;
            for (i64 N = B->klo[J]; N <= B->khi[J]; ++N) {
                if (A->c[N] == 0 || !sky_has(A, K, N)) continue;
                interior = interior && N >= own_lo && N < own_hi;
                MmSeg sg{sky_start(A, K, N), A->S[N], sky_start(B, N, J), B->S[J], A->c[N]};
                aligned = aligned && !(sg.astart & 1) && !(sg.lda & 1) && !(sg.bstart & 1) && !(sg.ldb & 1);
                aligned4 = aligned4 && !(sg.astart & 3) && !(sg.lda & 3) && !(sg.bstart & 3) && !(sg.ldb & 3);
                full128 = full128 && sg.kc % 32 == 0 && cb.rK % 128 == 0 && cb.cJ % 128 == 0;
                flops += 2.0 * (double)cb.rK * (double)cb.cJ * (double)sg.kc;
                segs.push_back(sg);
            }
            cb.seg_end = (int32_t)segs.size();
            const int32_t ci = (int32_t)cblks.size();
            cblks.push_back(cb);
            for (int t = 0; t < 2; ++t)
                for (int tj = 0; tj * tsz[t] < cb.cJ; ++tj)
                    for (int ti = 0; ti * tsz[t] < cb.rK; ++ti) (interior ? tiles[t] : btiles[t]).push_back(MmTile{ci, ti, tj, 0});
        }
    }
    int ninterior[2];
    for (int t = 0; t < 2; ++t) {
        ninterior[t] = (int)tiles[t].size();
        tiles[t].insert(tiles[t].end(), btiles[t].begin(), btiles[t].end());
    }
    if (segs.size() > (size_t)INT32_MAX / 2 || tiles[1].size() > (size_t)INT32_MAX / 2)
        return bbm_fail(h, BBM_ERR_UNSUPPORTED, "product plan too large");
    auto* pl = new (std::nothrow) bbm_sky_desc::MmPlan();
    if (!pl) return bbm_fail(h, BBM_ERR_ALLOC, "out of host memory");
    pl->aligned = aligned; pl->aligned4 = aligned4; pl->flops = flops; pl->full128 = full128 && aligned;
    pl->ninterior[0] = ninterior[0]; pl->ninterior[1] = ninterior[1];
    int32_t rc = upload(h, cblks, &pl->d_cblks);
    if (rc == BBM_OK) rc = upload(h, segs, &pl->d_segs);
    for (int t = 0; t < 2 && rc == BBM_OK; ++t) { rc = upload(h, tiles[t], &pl->d_tiles[t]); pl->ntiles[t] = (int)tiles[t].size(); }
    if (rc != BBM_OK) { free_plan(pl); return rc; }
    *out = pl;
    return BBM_OK;
}

// phase 0: the whole product; 1: only the tiles that multiply owned block rows [own_lo,own_hi) of B; 2: the others
// (sharded product: phase 1 runs while the halo rows of B are in flight, phase 2 behind their arrival)
template <typename T>
int32_t sky_matmul_impl(bbm_handle_t h, bbm_sky_desc_t A, bbm_sky_desc_t B, bbm_sky_desc_t Cd, T alpha, const T* dA,
                        const T* dB, T beta, T* dC, int phase = 0, i64 own_lo = 0, i64 own_hi = INT64_MAX) {
    BBM_REQUIRE_HANDLE(h);
    if (!sky_valid(A) || !sky_valid(B) || !sky_valid(Cd)) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BlockSkylineMatrix descriptor");
    BBM_CUDA(h, cudaSetDevice(h->device));
    auto key = std::make_pair(A->serial, B->serial);
    auto it = Cd->plans.find(key);
    if (it == Cd->plans.end()) {
        bbm_sky_desc::MmPlan* pl = nullptr;
        int32_t rc = build_mm_plan(h, A, B, Cd, &pl, own_lo, own_hi);
        if (rc != BBM_OK) return rc;
        it = Cd->plans.emplace(key, pl).first;
    }
    const bbm_sky_desc::MmPlan* pl = it->second;
    if (pl->ntiles[1] == 0) return BBM_OK;
    int t0[2], nt[2];
    for (int t = 0; t < 2; ++t) {
        t0[t] = phase == 2 ? pl->ninterior[t] : 0;
        nt[t] = phase == 0 ? pl->ntiles[t] : (phase == 1 ? pl->ninterior[t] : pl->ntiles[t] - pl->ninterior[t]);
    }
    if (nt[1] == 0) return BBM_OK;
    if (!dC || (A->numel > 0 && !dA) || (B->numel > 0 && !dB)) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    if (dC == dA || dC == dB) return bbm_fail(h, BBM_ERR_ARG, "mul!: destination must not alias an operand");
    MmParams<T> p{dA, dB, dC, pl->d_cblks, pl->d_segs, nullptr, alpha, beta};
    const i64 force = bbm_opt(h, "sky.mm_kernel", 0);
    cudaError_t e = cudaSuccess;
    if (sizeof(T) == 8 && force != 2) {
        const bool al = pl->aligned && (reinterpret_cast<uintptr_t>(dA) % 16 == 0) && (reinterpret_cast<uintptr_t>(dB) % 16 == 0);
        MmParams<double> pd{(const double*)dA, (const double*)dB, (double*)dC, pl->d_cblks, pl->d_segs, pl->d_tiles[0] + t0[0], (double)alpha, (double)beta};
        auto go = [&](auto kern, size_t smem) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) { kern<<<nt[0], 256, smem, h->stream>>>(pd); e = cudaGetLastError(); }
        };
        // k-chunk depth 32 halves the per-chunk barrier/issue bubble of the DMMA pipe (207 KB of smem)
        const bool bk32 = bbm_opt(h, "sky.mm_bk", 32) >= 32 && h->smem_optin >= mm_smem_bytes<128, 128, 32>();
        // warp-specialised bulk-copy pipeline for full tiles: opt-in ("sky.mm_ws" = 1) -- measured SLOWER than the cp.async kernel
        // on cfg4 (56.8 vs 54.0 ms): 160 bulk copies of 256 B - 1 KB per chunk keep the copy engine busier than the DMMA pipe
        const bool ws = al && pl->full128 && bbm_opt(h, "sky.mm_ws", 0) != 0 && h->smem_optin >= mm_ws_smem_bytes<128, 128, 32, 3>();
        if (ws) {
            constexpr size_t smem = mm_ws_smem_bytes<128, 128, 32, 3>();
            e = cudaFuncSetAttribute(sky_dgemm_dmma_ws_kernel<128, 128, 32, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) { sky_dgemm_dmma_ws_kernel<128, 128, 32, 3><<<nt[0], 288, smem, h->stream>>>(pd); e = cudaGetLastError(); }
        } else if (al && bk32) go(sky_dgemm_dmma_kernel<128, 128, 64, 32, 32, true>, mm_smem_bytes<128, 128, 32>());
        else if (al) go(sky_dgemm_dmma_kernel<128, 128, 64, 32, 16, true>, mm_smem_bytes<128, 128, 16>());
        else go(sky_dgemm_dmma_kernel<128, 128, 64, 32, 16, false>, mm_smem_bytes<128, 128, 16>());
    } else if (sizeof(T) == 4 && force != 2 && bbm_opt(h, "sky.mm_f32_kernel", 0) == 0) {
        MmParams<float> pf{(const float*)dA, (const float*)dB, (float*)dC, pl->d_cblks, pl->d_segs, pl->d_tiles[0] + t0[0], (float)alpha, (float)beta};
        const bool al = pl->aligned4 && (reinterpret_cast<uintptr_t>(dA) % 16 == 0) && (reinterpret_cast<uintptr_t>(dB) % 16 == 0);
        auto go = [&](auto kern) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSgSmem);
            if (e == cudaSuccess) { kern<<<nt[0], 256, kSgSmem, h->stream>>>(pf); e = cudaGetLastError(); }
        };
        if (al) go(sky_sgemm_simt_kernel<true>);
        else go(sky_sgemm_simt_kernel<false>);
    } else {
        p.tiles = pl->d_tiles[1] + t0[1];
        sky_gemm_simt_kernel<T><<<nt[1], 256, 0, h->stream>>>(p);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "block-banded product launch: %s", cudaGetErrorString(e));
    h->launches += 1;
    return BBM_OK;
}

// ------------------------------------------------------------------------------------------------
// block triangular solve  B <- T^-1 B,  T = Upper/LowerTriangular (unit or not) of a block-banded
// matrix with square diagonal blocks ([upstream] BlockArrays block substitution; bounds
// src/triblockbanded.jl:10-25, panel operand src/linalg.jl:93-115).
//   pass 0 (parallel):   invert every 32x32 diagonal sub-block of every diagonal block
//   per block K (sequential, as in the reference):
//       diag kernel   x_K <- T_KK^-1 b_K by blocked substitution over 32-row sub-blocks, each step
//                     a small product with the pre-inverted sub-block; right-hand sides split
//                     over CTAs, b_K kept in shared memory
//       panel update  b_KR <- b_KR - A[KR,K] x_K: the panel GEMV kernel on the contiguous
//                     off-diagonal panel (all right-hand sides at once, the matrix read once
//                     instead of once per column as on the CPU)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) sky_tri_invdiag_kernel(const T* __restrict__ data, const i64* __restrict__ sub, i64 nsub,
                                                              int upper, int unit, T* __restrict__ W) {
    __shared__ T Ts[4][32][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const i64 sb = (i64)blockIdx.x * 4 + warp;
    if (sb >= nsub) return;
    const i64 start = sub[3 * sb], ld = sub[3 * sb + 1];
    const int nb = (int)sub[3 * sb + 2];
    for (int j = 0; j < nb; ++j)
        if (lane < nb) Ts[warp][lane][j] = data[start + lane + (i64)j * ld];
    __syncwarp();
    // lane j computes column j of the inverse:  T x = e_j
    T x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = T(0);
    const int j = lane;
    if (j < nb) {
        if (upper) {
#pragma unroll
            for (int i = 31; i >= 0; --i) {
                if (i <= j && i < nb) {
                    T s = (i == j) ? T(1) : T(0);
#pragma unroll
                    for (int k = 31; k > 0; --k)
                        if (k > i && k <= j) s = fma(-Ts[warp][i][k], x[k], s);
                    x[i] = unit ? s : s / Ts[warp][i][i];
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (i >= j && i < nb) {
                    T s = (i == j) ? T(1) : T(0);
#pragma unroll
                    for (int k = 0; k < 31; ++k)
                        if (k < i && k >= j) s = fma(-Ts[warp][i][k], x[k], s);
                    x[i] = unit ? s : s / Ts[warp][i][i];
                }
            }
        }
    }
    T* w = W + sb * 1024;
#pragma unroll
    for (int i = 0; i < 32; ++i) w[i + 32 * j] = x[i];
}

template <typename T, int RC>
__global__ void __launch_bounds__(256) sky_tri_diag_kernel(const T* __restrict__ data, i64 dstart, i64 ld, int rK, int upper,
                                                           const T* __restrict__ W, T* __restrict__ Bk, i64 ldb, int nrhs) {
    extern __shared__ __align__(16) unsigned char tri_smem[];
    T* bs = reinterpret_cast<T*>(tri_smem);  // bs[q * rK + i]
    const int tid = threadIdx.x;
    const int q0 = blockIdx.x * RC;
    const int nq = min(RC, nrhs - q0);
    for (int idx = tid; idx < RC * rK; idx += 256) {
        const int q = idx / rK, i = idx % rK;
        bs[idx] = q < nq ? Bk[(i64)(q0 + q) * ldb + i] : T(0);
    }
    const int nsb = (rK + 31) / 32;
    const T* __restrict__ A = data + dstart;
    const int i = tid & 31, q = tid >> 5;
    // row i of the pre-inverted sub-block, kept in registers and prefetched one sub-step ahead
    T wreg[32];
    {
        const T* __restrict__ w = W + (i64)(upper ? nsb - 1 : 0) * 1024;
#pragma unroll
        for (int j = 0; j < 32; ++j) wreg[j] = (q < RC) ? w[i + 32 * j] : T(0);
    }
    __syncthreads();
    for (int s = 0; s < nsb; ++s) {
        const int ib = upper ? nsb - 1 - s : s;
        const int i0 = ib * 32, nb = min(32, rK - i0);
        // x_ib = inv(T_ib,ib) * b_ib   (entries of the inverse beyond nb are zero)
        T xv = T(0);
        if (q < RC) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < nb) xv = fma(wreg[j], bs[q * rK + i0 + j], xv);
        }
        __syncthreads();
        if (q < RC && i < nb) {
            bs[q * rK + i0 + i] = xv;
            if (q < nq) Bk[(i64)(q0 + q) * ldb + i0 + i] = xv;
        }
        if (s + 1 < nsb) {
            const T* __restrict__ w = W + (i64)(upper ? ib - 1 : ib + 1) * 1024;
#pragma unroll
            for (int j = 0; j < 32; ++j) wreg[j] = (q < RC) ? w[i + 32 * j] : T(0);
        }
        __syncthreads();
        // remaining rows of the diagonal block: b -= T[rows, ib] * x_ib; all 32 loads of a row are
        // issued before the first use so that one memory latency covers the whole sub-step
        const int rlo = upper ? 0 : i0 + nb, rhi = upper ? i0 : rK;
        for (int row = rlo + tid; row < rhi; row += 256) {
            const T* __restrict__ a = A + row + (i64)i0 * ld;
            T v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (j < nb) ? a[(i64)j * ld] : T(0);
            T acc[RC];
#pragma unroll
            for (int qq = 0; qq < RC; ++qq) acc[qq] = T(0);
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < nb) {
#pragma unroll
                    for (int qq = 0; qq < RC; ++qq) acc[qq] = fma(v[j], bs[qq * rK + i0 + j], acc[qq]);
                }
#pragma unroll
            for (int qq = 0; qq < RC; ++qq) bs[qq * rK + row] -= acc[qq];
        }
        __syncthreads();
    }
}

// pulls a byte range into L2 ahead of the sequential chain (the next block column's panel): one
// 4-byte ld.global.cg per 32-byte sector (prefetch.global.L2 hints were measured to be ineffective)
__global__ void __launch_bounds__(256) sky_l2_prefetch_kernel(const char* __restrict__ p, size_t bytes, unsigned* sink) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * 32;
    unsigned acc = 0;
    for (size_t off = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 32; off + 4 <= bytes; off += stride) {
        unsigned v;
        asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p + off));
        acc ^= v;
    }
    if (acc == 0x9E3779B9u && sink) *sink = acc;  // keeps the loads alive; practically never true
}

int32_t build_tri(bbm_handle_t h, bbm_sky_desc_t d, int upper) {
    bbm_sky_desc::Tri& t = d->tri[upper];
    if (t.built) return BBM_OK;
    const i64 N = d->Nr;
    std::vector<SkyEnt> ents(N);
    std::vector<SkyItem> items;
    std::vector<i64> sub;
    t.item_ptr.assign(N + 1, 0);
    t.sub_ptr.assign(N + 1, 0);
    for (i64 K = 0; K < N; ++K) {
        t.item_ptr[K] = (int)items.size();
        t.sub_ptr[K + 1] = t.sub_ptr[K] + (d->r[K] + 31) / 32;
        ents[K] = SkyEnt{0, 1, d->R[K], d->r[K]};
        if (d->r[K] == 0) continue;
        const i64 dstart = sky_start(d, K, K);
        for (i64 i0 = 0; i0 < d->r[K]; i0 += 32) {
            sub.push_back(dstart + i0 + i0 * d->S[K]);
            sub.push_back(d->S[K]);
            sub.push_back(std::min<i64>(32, d->r[K] - i0));
        }
        // off-diagonal panel of block column K: rows of block rows klo..K-1 (upper) / K+1..khi (lower)
        i64 prow0, prows, yrow0;
        if (upper) { prow0 = 0; prows = d->R[K] - d->R[d->klo[K]]; yrow0 = d->R[d->klo[K]]; }
        else { prow0 = d->R[K + 1] - d->R[d->klo[K]]; prows = d->R[d->khi[K] + 1] - d->R[K + 1]; yrow0 = d->R[K + 1]; }
        ents[K] = SkyEnt{d->off[K] + prow0, d->S[K], d->R[K], d->r[K]};
        for (i64 r0 = 0; r0 < prows; r0 += kGemvRows)
            items.push_back(SkyItem{yrow0 + r0, r0, (int32_t)std::min<i64>(kGemvRows, prows - r0), (int32_t)K, (int32_t)K + 1, 0});
    }
    t.item_ptr[N] = (int)items.size();
    t.nsub = (i64)sub.size() / 3;
    int32_t rc = upload(h, ents, &t.d_ents);
    if (rc == BBM_OK) rc = upload(h, items, &t.d_items);
    if (rc == BBM_OK) rc = upload(h, sub, &t.d_sub);
    if (rc != BBM_OK) return rc;
    t.built = true;
    return BBM_OK;
}

// ------------------------------------------------------------------------------------------------
// GEMM-based block triangular solve (Float64, >= 16 right-hand sides).
// The blocked substitution above is bound by its dependency chain (16 sub-steps per 512-row block, each a
// latency-exposed shared-memory product).  Here every diagonal block is inverted explicitly up front --
// 32x32 sub-blocks by substitution, then doubling: for T = [T11 T12; 0 T22], inv(T)12 = -inv(T11) T12 inv(T22),
// two batched DMMA products per level over ALL diagonal blocks at once -- and each block step becomes two
// tensor-core products with split-k (atomic) accumulation:
//     X_K  = inv(T_KK) * B_K                 (only the non-zero triangle of the inverse is multiplied)
//     B_KR = B_KR - A[KR,K] * X_K            (the off-diagonal panel of src/linalg.jl:93-115)
// Rounding differs from substitution by O(cond(T_KK) eps); the tolerance tests run on the reference's
// well-conditioned inputs.  "trsm.kernel" = 2 keeps the substitution path.
// ------------------------------------------------------------------------------------------------
struct MmBatchHost {
    std::vector<MmCBlk> cblks;
    std::vector<MmSeg> segs;
    std::vector<MmTile> tiles;
    bool aligned = true;
    int tile = 64;   // output tile edge: 64 (4-warp kernel, split-k chain links) or 128 (8-warp kernel, the large inversion levels)
    // one C block of rows x cols at (cstart, ldc) with a single product segment, cut into tile x tile pieces
    void add(i64 cstart, i64 ldc, i64 rows, i64 cols, const MmSeg& sg) {
        if (rows <= 0 || cols <= 0 || sg.kc <= 0) return;
        MmCBlk cb{cstart, ldc, (int32_t)rows, (int32_t)cols, (int32_t)segs.size(), (int32_t)segs.size() + 1};
        aligned = aligned && !(sg.astart & 1) && !(sg.lda & 1) && !(sg.bstart & 1) && !(sg.ldb & 1);
        segs.push_back(sg);
        const int32_t ci = (int32_t)cblks.size();
        cblks.push_back(cb);
        for (int tj = 0; tj * tile < cols; ++tj)
            for (int ti = 0; ti * tile < rows; ++ti) tiles.push_back(MmTile{ci, ti, tj, 0});
    }
    // a C block with an empty product: it only gets its beta scaling (or zeros)
    void add_empty(i64 cstart, i64 ldc, i64 rows, i64 cols) {
        if (rows <= 0 || cols <= 0) return;
        const int32_t ci = (int32_t)cblks.size();
        cblks.push_back(MmCBlk{cstart, ldc, (int32_t)rows, (int32_t)cols, (int32_t)segs.size(), (int32_t)segs.size()});
        for (int tj = 0; tj * tile < cols; ++tj)
            for (int ti = 0; ti * tile < rows; ++ti) tiles.push_back(MmTile{ci, ti, tj, 0});
    }
    // C (rows x cols) = A (rows x kdim) * B (kdim x cols) where one factor is TRIANGULAR (its zero triangle is skipped):
    // tri = 1: B upper triangular (output columns [c0, c0+nc) need k < c0+nc); 2: B lower triangular (k >= c0);
    //       3: A upper triangular (output rows [r0, r0+nr) need k >= r0);       4: A lower triangular (k < r0+nr).
    // Every tile x tile piece of C becomes a block of its own with its own trimmed k range.
    void add_tri(i64 cstart, i64 ldc, i64 rows, i64 cols, i64 astart, i64 lda, i64 bstart, i64 ldb, i64 kdim, int tri) {
        for (i64 c0 = 0; c0 < cols; c0 += tile)
            for (i64 r0 = 0; r0 < rows; r0 += tile) {
                const i64 nr = std::min<i64>(tile, rows - r0), nc = std::min<i64>(tile, cols - c0);
                i64 klo = 0, khi = kdim;
                if (tri == 1) khi = std::min(kdim, c0 + nc);
                else if (tri == 2) klo = std::min(kdim, c0);
                else if (tri == 3) klo = std::min(kdim, r0);
                else if (tri == 4) khi = std::min(kdim, r0 + nr);
                klo &= ~i64(1);   // keep 16-byte alignment of the operands' k offsets
                add(cstart + r0 + c0 * ldc, ldc, nr, nc, MmSeg{astart + r0 + klo * lda, lda, bstart + klo + c0 * ldb, ldb, khi - klo});
            }
    }
};

void free_batch(bbm_sky_desc::Tri::Batch& b) {
    if (b.cblks) cudaFree(b.cblks);
    if (b.segs) cudaFree(b.segs);
    if (b.tiles) cudaFree(b.tiles);
    b = bbm_sky_desc::Tri::Batch();
}

int32_t upload_batch(bbm_handle_t h, const MmBatchHost& hb, bbm_sky_desc::Tri::Batch* out) {
    free_batch(*out);
    MmCBlk* c = nullptr; MmSeg* sgs = nullptr; MmTile* t = nullptr;
    int32_t rc = upload(h, hb.cblks, &c);
    if (rc == BBM_OK) rc = upload(h, hb.segs, &sgs);
    if (rc == BBM_OK) rc = upload(h, hb.tiles, &t);
    out->cblks = c; out->segs = sgs; out->tiles = t;
    out->ntiles = (int)hb.tiles.size();
    out->aligned = hb.aligned;
    out->tile = hb.tile;
    return rc;
}

// C (+)= alpha * A * B over the tiles [t0, t0+nt) of a batch; 64x64 DMMA tiles, 4 warps
int32_t launch_mm64(bbm_handle_t h, const bbm_sky_desc::Tri::Batch& b, int t0, int nt, const double* A, const double* Bm,
                    double* Cm, double alpha, double beta, bool atomic, bool set_attr = true, bool pdl = false, cudaStream_t st = nullptr) {
    if (nt <= 0) return BBM_OK;
    if (!st) st = h->stream;
    MmParams<double> p{A, Bm, Cm, (const MmCBlk*)b.cblks, (const MmSeg*)b.segs, (const MmTile*)b.tiles + t0, alpha, beta};
    const bool al = b.aligned && (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (reinterpret_cast<uintptr_t>(Bm) % 16 == 0);
    if (b.tile == 128) {
        // the large levels of the inversion: the 8-warp 128 x 128 kernels of the block-banded product
        if (atomic) return bbm_fail(h, BBM_ERR_ARG, "internal: 128-tiles are not split");
        cudaError_t e2 = cudaSuccess;
        auto run = [&](auto kern, size_t sm) {
            e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            if (e2 == cudaSuccess) { kern<<<nt, 256, sm, h->stream>>>(p); e2 = cudaGetLastError(); }
        };
        if (al && h->smem_optin >= mm_smem_bytes<128, 128, 32>()) run(sky_dgemm_dmma_kernel<128, 128, 64, 32, 32, true>, mm_smem_bytes<128, 128, 32>());
        else if (al) run(sky_dgemm_dmma_kernel<128, 128, 64, 32, 16, true>, mm_smem_bytes<128, 128, 16>());
        else run(sky_dgemm_dmma_kernel<128, 128, 64, 32, 16, false>, mm_smem_bytes<128, 128, 16>());
        if (e2 != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "triangular-solve inversion product: %s", cudaGetErrorString(e2));
        h->launches += 1;
        return BBM_OK;
    }
    constexpr size_t smem = mm_smem_bytes<64, 64, 16, 4>();  // 4 stages: a 64-deep k-slice is requested at once
    cudaError_t e = cudaSuccess;
    const char* where = "";
    auto go = [&](auto kern) {
        if (set_attr) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            where = "cudaFuncSetAttribute";
        }
        if (e == cudaSuccess) {
            // programmatic dependent launch: the kernel's launch latency and prologue overlap the predecessor's
            // tail; griddepcontrol.wait inside the kernel still orders all operand accesses after it
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)nt); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
            e = cudaLaunchKernelEx(&cfg, kern, p);
            where = "launch";
        }
    };
    if (al && atomic) go(sky_dgemm_dmma_kernel<64, 64, 32, 32, 16, true, true, 4>);
    else if (al) go(sky_dgemm_dmma_kernel<64, 64, 32, 32, 16, true, false, 4>);
    else if (atomic) go(sky_dgemm_dmma_kernel<64, 64, 32, 32, 16, false, true, 4>);
    else go(sky_dgemm_dmma_kernel<64, 64, 32, 32, 16, false, false, 4>);
    if (e != cudaSuccess)
        return bbm_fail(h, BBM_ERR_CUDA, "triangular-solve product %s (tiles %d..+%d of %d, aligned %d, atomic %d): %s", where, t0, nt,
                        b.ntiles, (int)al, (int)atomic, cudaGetErrorString(e));
    h->launches += 1;
    return BBM_OK;
}

int32_t build_tri_inverse_plan(bbm_handle_t h, bbm_sky_desc_t d, int upper) {
    bbm_sky_desc::Tri& t = d->tri[upper];
    if (t.inv_built) return BBM_OK;
    const i64 N = d->Nr;
    t.woff.assign(N + 1, 0); t.wld.assign(N, 2);
    i64 maxr = 0;
    for (i64 K = 0; K < N; ++K) {
        t.wld[K] = std::max<i64>(2, (d->r[K] + 1) & ~i64(1));
        t.woff[K + 1] = t.woff[K] + t.wld[K] * ((d->r[K] + 1) & ~i64(1));
        maxr = std::max(maxr, d->r[K]);
    }
    t.wtotal = t.woff[N];
    {
        std::vector<i64> place;
        for (i64 K = 0; K < N; ++K)
            for (i64 i0 = 0; i0 < d->r[K]; i0 += 32) {
                place.push_back(t.woff[K] + i0 + i0 * t.wld[K]);
                place.push_back(t.wld[K]);
                place.push_back(std::min<i64>(32, d->r[K] - i0));
            }
        if (t.d_place) { cudaFree(t.d_place); t.d_place = nullptr; }
        int32_t rcp = upload(h, place, &t.d_place);
        if (rcp != BBM_OK) return rcp;
        std::vector<i64> zero;
        for (i64 K = 0; K < N; ++K) {
            const i64 rK = d->r[K], wl = t.wld[K];
            const i64 rpad = (rK + 1) & ~i64(1);   // the storage of an odd block has one padding row / column
            for (i64 j0 = 0; j0 < rpad; j0 += 32) {
                const i64 nc = std::min<i64>(32, rpad - j0);
                // upper triangle: rows below the strip's diagonal sub-block; lower triangle: rows above it
                const i64 i0 = upper ? j0 + nc : 0, nr = upper ? rpad - (j0 + nc) : j0;
                if (nr > 0) { zero.push_back(t.woff[K] + i0 + j0 * wl); zero.push_back(wl); zero.push_back(nr); zero.push_back(nc); }
            }
            if (rpad > rK) {   // the padding row and column themselves (read by 2-element vector loads at the edge)
                zero.push_back(t.woff[K] + rK); zero.push_back(wl); zero.push_back(1); zero.push_back(rpad);
                zero.push_back(t.woff[K] + rK * wl); zero.push_back(wl); zero.push_back(rpad); zero.push_back(1);
            }
        }
        if (t.d_zero) { cudaFree(t.d_zero); t.d_zero = nullptr; }
        t.nzero = (i64)zero.size() / 4;
        rcp = upload(h, zero, &t.d_zero);
        if (rcp != BBM_OK) return rcp;
    }
    for (auto& b : t.inv_levels) free_batch(b);
    t.inv_levels.clear();
    for (i64 sz = 32; sz < maxr; sz *= 2) {
        MmBatchHost g1, g2;  // g1: T = offdiag * X22 (upper) / offdiag * X11 (lower);  g2: X12 = -X11 * T / X21 = -X22 * T
        // the inverses of the sub-blocks are triangular: their zero triangle is skipped tile by tile (37 % of the flops of the
        // largest level), and the two large levels use 128 x 128 tiles (the 8-warp kernel: 31 instead of 23 TFLOP/s)
        g1.tile = g2.tile = sz >= 128 ? 128 : 64;
        for (i64 K = 0; K < N; ++K) {
            const i64 rK = d->r[K];
            if (rK <= sz) continue;
            const i64 ds = sky_start(d, K, K), ld = d->S[K], wo = t.woff[K], wl = t.wld[K];
            for (i64 base = 0; base + sz < rK; base += 2 * sz) {
                const i64 s1 = sz, s2 = std::min<i64>(sz, rK - base - sz), mid = base + s1;
                if (upper) {
                    // T12 (s1 x s2) = U12 * X22 (upper triangular) ;  X12 = -X11 (upper triangular) * T12
                    g1.add_tri(wo + base + mid * wl, wl, s1, s2, ds + base + mid * ld, ld, wo + mid + mid * wl, wl, s2, 1);
                    g2.add_tri(wo + base + mid * wl, wl, s1, s2, wo + base + base * wl, wl, wo + base + mid * wl, wl, s1, 3);
                } else {
                    // T21 (s2 x s1) = L21 * X11 (lower triangular) ;  X21 = -X22 (lower triangular) * T21
                    g1.add_tri(wo + mid + base * wl, wl, s2, s1, ds + mid + base * ld, ld, wo + base + base * wl, wl, s1, 2);
                    g2.add_tri(wo + mid + base * wl, wl, s2, s1, wo + mid + mid * wl, wl, wo + mid + base * wl, wl, s2, 4);
                }
            }
        }
        t.inv_levels.emplace_back(); t.inv_levels.emplace_back();
        int32_t rc = upload_batch(h, g1, &t.inv_levels[t.inv_levels.size() - 2]);
        if (rc == BBM_OK) rc = upload_batch(h, g2, &t.inv_levels[t.inv_levels.size() - 1]);
        if (rc != BBM_OK) return rc;
    }
    t.inv_built = true;
    return BBM_OK;
}

// scatter the inverted 32x32 diagonal sub-blocks (1024 entries each) into the dense inverse blocks
__global__ void __launch_bounds__(256) sky_tri_scatter_kernel(const double* __restrict__ W32, const i64* __restrict__ place, i64 nsub,
                                                              double* __restrict__ W) {
    const i64 sb = blockIdx.x;
    if (sb >= nsub) return;
    const i64 dst = place[3 * sb], ld = place[3 * sb + 1];
    const int nb = (int)place[3 * sb + 2];
    for (int idx = threadIdx.x; idx < 1024; idx += 256) {
        const int i = idx & 31, j = idx >> 5;
        if (i < nb && j < nb) W[dst + i + (i64)j * ld] = W32[sb * 1024 + idx];
    }
}

// Zero the zero side of the triangle of every inverse block (one rectangle per 32-column strip): the doubling levels and the
// chain's products read the inverses as dense operands, and every other entry of W is written by the scatter above or by a
// doubling level -- half the bytes of the memset of all of W that this replaces.
__global__ void __launch_bounds__(256) sky_tri_zero_kernel(const i64* __restrict__ zt, i64 nz, double* __restrict__ W) {
    const i64 z = blockIdx.x;
    if (z >= nz) return;
    const i64 dst = zt[4 * z], ld = zt[4 * z + 1];
    const int nr = (int)zt[4 * z + 2], nc = (int)zt[4 * z + 3];
    for (int idx = threadIdx.x; idx < nr * nc; idx += 256) {
        const int i = idx % nr, j = idx / nr;
        W[dst + i + (i64)j * ld] = 0.0;
    }
}

int32_t build_tri_step_plan(bbm_handle_t h, bbm_sky_desc_t d, int upper, i64 ldb, i64 ldx, i64 nrhs) {
    bbm_sky_desc::Tri& t = d->tri[upper];
    if (t.step_ldb == ldb && t.step_nrhs == nrhs && t.step_ldx == ldx && t.stepA.tiles) return BBM_OK;
    const i64 N = d->Nr;
    MmBatchHost a, b;
    t.ptrA.assign(N + 1, 0); t.ptrB.assign(N + 1, 0);
    const i64 ksplit = 64;  // k-slice per CTA: latency-bound products, so many short CTAs beat few long ones
    for (i64 K = 0; K < N; ++K) {
        t.ptrA[K] = (int)a.tiles.size();
        t.ptrB[K] = (int)b.tiles.size();
        const i64 rK = d->r[K];
        if (rK == 0) continue;
        const i64 wo = t.woff[K], wl = t.wld[K];
        // X[R_K + rows of tile ti] += inv(T)[rows, k-part] * B[R_K + k-part]   (skip the zero triangle)
        for (i64 i0 = 0; i0 < rK; i0 += 64) {
            const i64 nr = std::min<i64>(64, rK - i0);
            const i64 klo = upper ? i0 : 0, khi = upper ? rK : std::min<i64>(rK, i0 + 64);
            for (i64 k0 = klo; k0 < khi; k0 += ksplit) {
                const i64 kc = std::min<i64>(ksplit, khi - k0);
                a.add(d->R[K] + i0, ldx, nr, nrhs, MmSeg{wo + i0 + k0 * wl, wl, d->R[K] + k0, ldb, kc});
            }
        }
        // B[panel rows] -= A[panel, K] * X[R_K ...]
        i64 prow0, prows, yrow0;
        if (upper) { prow0 = 0; prows = d->R[K] - d->R[d->klo[K]]; yrow0 = d->R[d->klo[K]]; }
        else { prow0 = d->R[K + 1] - d->R[d->klo[K]]; prows = d->R[d->khi[K] + 1] - d->R[K + 1]; yrow0 = d->R[K + 1]; }
        for (i64 k0 = 0; k0 < rK && prows > 0; k0 += ksplit) {
            const i64 kc = std::min<i64>(ksplit, rK - k0);
            b.add(yrow0, ldb, prows, nrhs, MmSeg{d->off[K] + prow0 + k0 * d->S[K], d->S[K], d->R[K] + k0, ldx, kc});
        }
    }
    t.ptrA[N] = (int)a.tiles.size();
    t.ptrB[N] = (int)b.tiles.size();
    int32_t rc = upload_batch(h, a, &t.stepA);
    if (rc == BBM_OK) rc = upload_batch(h, b, &t.stepB);
    if (rc != BBM_OK) return rc;
    t.step_ldb = ldb; t.step_nrhs = nrhs; t.step_ldx = ldx;
    return BBM_OK;
}

int32_t build_tri_flow_plan(bbm_handle_t h, bbm_sky_desc_t d, int upper, i64 ldb, i64 ldx, i64 nrhs);
int32_t launch_trsm_flow(bbm_handle_t h, bbm_sky_desc_t d, int upper, const double* d_data, const double* W, double* d_B, double* X);
int32_t tri_condition(bbm_handle_t h, bbm_sky_desc_t d, int upper, int unit, const double* d_data, const double* W, double* cond_max);

// returns BBM_OK and *fallback = true (nothing written to B) when the condition guard rejects the inverse-based paths
int32_t sky_trsm_gemm(bbm_handle_t h, bbm_sky_desc_t d, int upper, int unit, const double* d_data, double* d_B, i64 ldb, i64 nrhs, bool* fallback) {
    *fallback = false;
    bbm_sky_desc::Tri& t = d->tri[upper];
    int32_t rc = build_tri_inverse_plan(h, d, upper);
    if (rc != BBM_OK) return rc;
    const i64 ldx = std::max<i64>(2, (d->m + 1) & ~i64(1));
    rc = build_tri_step_plan(h, d, upper, ldb, ldx, nrhs);
    if (rc != BBM_OK) return rc;
    // workspaces: W (inverses), Tw (level temporaries, same layout), X (solutions), W32 (32x32 inverses)
    rc = bbm_ws_reserve(h, &h->ws_x, &h->ws_x_bytes, (size_t)t.nsub * 1024 * sizeof(double));
    if (rc == BBM_OK) rc = bbm_ws_reserve(h, &h->ws_y, &h->ws_y_bytes, (size_t)ldx * nrhs * sizeof(double));
    if (rc == BBM_OK) rc = bbm_ws_reserve(h, &h->ws_z, &h->ws_z_bytes, (size_t)t.wtotal * sizeof(double));
    if (rc == BBM_OK) rc = bbm_ws_reserve(h, &h->ws_w, &h->ws_w_bytes, (size_t)t.wtotal * sizeof(double));
    if (rc != BBM_OK) return rc;
    double* W32 = static_cast<double*>(h->ws_x);
    double* X = static_cast<double*>(h->ws_y);
    double* W = static_cast<double*>(h->ws_z);
    double* Tw = static_cast<double*>(h->ws_w);
    if (t.nzero > 0) {
        sky_tri_zero_kernel<<<(unsigned)t.nzero, 256, 0, h->stream>>>(t.d_zero, t.nzero, W);
        h->launches += 1;
    }
    BBM_CUDA(h, cudaMemsetAsync(X, 0, (size_t)ldx * nrhs * sizeof(double), h->stream));
    // level 0: 32x32 inverses by substitution, scattered onto the diagonals of the dense inverse blocks
    sky_tri_invdiag_kernel<double><<<(unsigned)((t.nsub + 3) / 4), 128, 0, h->stream>>>(d_data, t.d_sub, t.nsub, upper, unit ? 1 : 0, W32);
    sky_tri_scatter_kernel<<<(unsigned)t.nsub, 256, 0, h->stream>>>(W32, t.d_place, t.nsub, W);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "sky_tri_invdiag launch: %s", cudaGetErrorString(e));
    h->launches += 2;
    // doubling levels, all diagonal blocks at once
    for (size_t lv = 0; lv + 1 < t.inv_levels.size(); lv += 2) {
        rc = launch_mm64(h, t.inv_levels[lv], 0, t.inv_levels[lv].ntiles, d_data, W, Tw, 1.0, 0.0, false);
        if (rc == BBM_OK) rc = launch_mm64(h, t.inv_levels[lv + 1], 0, t.inv_levels[lv + 1].ntiles, W, Tw, W, -1.0, 0.0, false);
        if (rc != BBM_OK) return rc;
    }
    // Condition guard.  X_K = inv(T_KK) B_K is not backward stable: its error grows like cond(T_KK) eps, where the
    // reference's substitution (trsv) stays at the n eps level.  ||T_KK||_inf ||inv(T_KK)||_inf of every diagonal block is
    // read back (one synchronisation per solve); beyond "trsm.cond_max" (default 8 x the longest row of T: the inverse path
    // then still meets 64 eps nnz) the call falls back to the substitution kernels.  "trsm.cond_check" = 0 skips it.
    if (bbm_opt(h, "trsm.cond_check", 1) != 0) {
        double cmax = 0.0;
        rc = tri_condition(h, d, upper, unit, d_data, W, &cmax);
        if (rc != BBM_OK) return rc;
        i64 longest = 1;
        for (i64 K = 0; K < d->Nr; ++K) {
            i64 len = 0;   // entries of block row K of T (diagonal block + off-diagonal blocks on the triangle's side)
            for (i64 J = 0; J < d->Nc; ++J) if (sky_has(d, K, J) && (upper ? J >= K : J <= K)) len += d->c[J];
            longest = std::max(longest, len);
            if (d->Nr > 4096) break;   // (constant-bandwidth structures: the first rows are representative enough)
        }
        const double limit = (double)bbm_opt(h, "trsm.cond_max", 8 * longest);
        h->last_cond = cmax;
        h->opts["trsm.last_cond_x1000"] = cmax < 9.0e15 ? (i64)(cmax * 1000.0) : INT64_MAX;   // diagnostics (bbm_get_option)
        if (!(cmax <= limit)) { *fallback = true; h->opts["trsm.last_path"] = 3; return BBM_OK; }
    }
    // the data-flow kernel: the whole chain in one persistent launch ("trsm.kernel" = 3 keeps the launch chain below)
    if (bbm_opt(h, "trsm.kernel", 0) != 3) {
        rc = build_tri_flow_plan(h, d, upper, ldb, ldx, nrhs);
        if (rc != BBM_OK) return rc;
        if (t.flow_ok && (reinterpret_cast<uintptr_t>(d_data) % 16 == 0) && (reinterpret_cast<uintptr_t>(d_B) % 16 == 0)) {
            rc = launch_trsm_flow(h, d, upper, d_data, W, d_B, X);
            if (rc != BBM_OK) return rc;
            h->opts["trsm.last_path"] = 1;
            BBM_CUDA(h, cudaMemcpy2DAsync(d_B, (size_t)ldb * sizeof(double), X, (size_t)ldx * sizeof(double), (size_t)d->m * sizeof(double),
                                          (size_t)nrhs, cudaMemcpyDeviceToDevice, h->stream));
            return BBM_OK;
        }
    }
    // the sequential chain: two split-k tensor-core products per block; the next block column's panel and
    // inverse are pulled into L2 on a low-priority side stream one step ahead
    // (measured on cfg5: 27.1 us per block step without, 27.8 us with the prefetch -- off by default here)
    const bool prefetch = bbm_opt(h, "trsm.prefetch", 0) != 0 && d->Nr > 2;
    if (prefetch && !h->side_stream) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        BBM_CUDA(h, cudaStreamCreateWithPriority(&h->side_stream, cudaStreamNonBlocking, lo));
        for (int i = 0; i < 2; ++i) BBM_CUDA(h, cudaEventCreateWithFlags(&h->side_ev[i], cudaEventDisableTiming));
    }
    if (prefetch) {
        BBM_CUDA(h, cudaEventRecord(h->side_ev[0], h->stream));
        BBM_CUDA(h, cudaStreamWaitEvent(h->side_stream, h->side_ev[0], 0));
    }
    h->opts["trsm.last_path"] = 2;
    for (i64 s = 0; s < d->Nr; ++s) {
        const i64 K = upper ? d->Nr - 1 - s : s;
        if (prefetch && s + 1 < d->Nr) {
            const i64 Kn = upper ? K - 1 : K + 1;
            const size_t pb = (size_t)(d->off[Kn + 1] - d->off[Kn]) * sizeof(double);
            const size_t wb = (size_t)(t.woff[Kn + 1] - t.woff[Kn]) * sizeof(double);
            if (pb > 0) {
                sky_l2_prefetch_kernel<<<(unsigned)std::min<size_t>(592, (pb / 32 + 255) / 256 + 1), 256, 0, h->side_stream>>>(
                    reinterpret_cast<const char*>(d_data + d->off[Kn]), pb, nullptr);
                h->launches += 1;
            }
            if (wb > 0) {
                sky_l2_prefetch_kernel<<<(unsigned)std::min<size_t>(592, (wb / 32 + 255) / 256 + 1), 256, 0, h->side_stream>>>(
                    reinterpret_cast<const char*>(W + t.woff[Kn]), wb, nullptr);
                h->launches += 1;
            }
        }
        // (the opt-in shared-memory size of the kernel variant only has to be set once per call)
        const bool pdl = bbm_opt(h, "trsm.pdl", 1) != 0;
        rc = launch_mm64(h, t.stepA, t.ptrA[K], t.ptrA[K + 1] - t.ptrA[K], W, d_B, X, 1.0, 1.0, true, s < 2, pdl);
        if (rc == BBM_OK) rc = launch_mm64(h, t.stepB, t.ptrB[K], t.ptrB[K + 1] - t.ptrB[K], d_data, X, d_B, -1.0, 1.0, true, s < 2, pdl);
        if (rc != BBM_OK) return rc;
        if (prefetch) {
            BBM_CUDA(h, cudaEventRecord(h->side_ev[s & 1], h->stream));
            BBM_CUDA(h, cudaStreamWaitEvent(h->side_stream, h->side_ev[s & 1], 0));
        }
    }
    if (prefetch) {
        BBM_CUDA(h, cudaEventRecord(h->side_ev[0], h->side_stream));
        BBM_CUDA(h, cudaStreamWaitEvent(h->stream, h->side_ev[0], 0));
    }
    BBM_CUDA(h, cudaMemcpy2DAsync(d_B, (size_t)ldb * sizeof(double), X, (size_t)ldx * sizeof(double), (size_t)d->m * sizeof(double),
                                  (size_t)nrhs, cudaMemcpyDeviceToDevice, h->stream));
    return BBM_OK;
}


// ------------------------------------------------------------------------------------------------
// Data-flow block triangular solve (Float64, >= 16 right-hand sides): the whole chain in ONE persistent kernel.
// The launch-per-link chain above spends most of a block step in launch gaps, cold prologues and HBM latency
// (22 us per step on cfg5 for ~3 us of tensor work).  Here every 64 x 32 x 64 piece of the two products of a step
//     X_K  += inv(T_KK)[i-tile, k-slice] * B_K[k-slice, j-tile]            ("X tile", waits until B_K is complete)
//     B_Kr -= A[Kr-tile, K][:, k-slice]  * X_K[k-slice, j-tile]            ("update tile", waits until X_K is complete)
// is a tile of one global list in dependency order.  Resident CTAs (cooperative launch: all of them are co-resident) take
// the tiles round-robin; each tile first prefetches its STATIC operand (a piece of the matrix or of an inverse -- neither
// depends on the chain) into shared memory, then waits for its completion counter in L2 (acquire), loads the 16 KB of
// right-hand-side data that do depend on the chain, multiplies on the FP64 tensor cores, accumulates with red.add and
// bumps the counter of its target (release).  Launch gaps and HBM latency leave the critical path; the two far updates
// of a step (2/3 of the flops) overlap the next steps' critical tiles instead of delaying them.  Summation order of the
// split-k partial sums is not deterministic ("trsm.kernel" = 3 keeps the launch chain, 2 the substitution kernels).
// ------------------------------------------------------------------------------------------------
struct FlowTile {
    i64 astart, lda;      // static operand: rows x kc, element (i,k) at astart + i + k*lda
    i64 bstart, ldb;      // dynamic operand: kc x cols, element (k,j) at bstart + k + j*ldb
    i64 cstart, ldc;      // destination tile, element (i,j) at cstart + i + j*ldc
    int32_t rows, cols, kc, kind;          // kind 0: X tile (inverse * B -> X, +), 1: update tile (matrix * X -> B, -)
    int32_t wait_ctr, wait_val, sig_ctr, pad;
};

constexpr int kFlowBM = 64, kFlowBN = 32;
constexpr int kFlowCtrStride = 32;   // ints between two completion counters: one 128-byte line per counter
constexpr int kFlowLDA = kFlowBM + 4;
template <int BK>
constexpr size_t flow_smem_bytes() { return ((size_t)BK * kFlowLDA + (size_t)kFlowBN * (BK + 4)) * sizeof(double); }

// Polling load: relaxed at gpu scope (served by L2, no L1 invalidation per iteration -- an acquire in the loop costs a CCTL.IVALL
// every time, 8.5 % of all stall samples in the first version); the acquire fence follows once, after the value has arrived.
__device__ __forceinline__ int ld_relaxed_gpu_s32(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int kFlowBK>
__global__ void __launch_bounds__(128) sky_trsm_flow_kernel(const FlowTile* __restrict__ tiles, i64 ntiles, const double* __restrict__ data,
                                                            const double* __restrict__ W, double* __restrict__ Bm, double* __restrict__ X,
                                                            int* __restrict__ ctr, int err_idx, int sleep_ns, int cs) {
    // cs: stride between counters in ints (32 = one 128-byte line each: waiting tiles of one step poll different L2 lines)
    constexpr int kFlowLDB = kFlowBK + 4;
    extern __shared__ __align__(16) unsigned char mm_smem[];
    double* As = reinterpret_cast<double*>(mm_smem);          // As[k][i]
    double* Bs = As + (size_t)kFlowBK * kFlowLDA;              // Bs[j][k]
    __shared__ int bad_flag;
    int* bad = &bad_flag;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm0 = (warp & 1) * 32, wn0 = (warp >> 1) * 16;
    const int ar = lane >> 2, ac = lane & 3;
    for (i64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const FlowTile ft = tiles[t];
        const double* __restrict__ Ap = ft.kind == 0 ? W : data;
        const double* __restrict__ Bp = ft.kind == 0 ? Bm : X;
        double* __restrict__ Cp = ft.kind == 0 ? X : Bm;
        // 1. the static operand (independent of the chain): prefetch while the predecessors still run
        for (int v = tid; v < (kFlowBM / 2) * kFlowBK; v += 128) {
            const int kk = v / (kFlowBM / 2), i = 2 * (v % (kFlowBM / 2));
            int nv = ft.rows - i; nv = nv < 0 ? 0 : (nv > 2 ? 2 : nv);
            const int bytes = kk < ft.kc ? 8 * nv : 0;
            cp_async16(As + kk * kFlowLDA + i, bytes ? Ap + ft.astart + i + (i64)kk * ft.lda : Ap, bytes);
        }
        cp_async_commit();
        // 2. wait until every contribution to the dynamic operand has landed (counter in L2)
        if (tid == 0) {
            int timed_out = 0;
            if (ft.wait_val > 0) {
                const int* wp = ctr + (i64)ft.wait_ctr * cs;
                if (ld_relaxed_gpu_s32(wp) < ft.wait_val) {
                    const long long t0 = clock64();
                    while (ld_relaxed_gpu_s32(wp) < ft.wait_val) {
                        if (sleep_ns > 0) __nanosleep(sleep_ns);
                        if (clock64() - t0 > 4000000000LL) { atomicExch(ctr + (i64)err_idx * cs, 1); timed_out = 1; break; }   // never hang the device
                    }
                }
                asm volatile("fence.acquire.gpu;" ::: "memory");
            }
            *bad = timed_out;   // a tile that gave up poisons its destination: everything downstream turns NaN, never silently wrong
        }
        __syncthreads();
        // 3. the dynamic operand (right-hand-side data written by red.add in L2: .cg loads bypass L1)
        for (int v = tid; v < kFlowBN * (kFlowBK / 2); v += 128) {
            const int j = v / (kFlowBK / 2), kk = 2 * (v % (kFlowBK / 2));
            int nv = ft.kc - kk; nv = nv < 0 ? 0 : (nv > 2 ? 2 : nv);
            const int bytes = j < ft.cols ? 8 * nv : 0;
            cp_async16(Bs + j * kFlowLDB + kk, bytes ? Bp + ft.bstart + kk + (i64)j * ft.ldb : Bp, bytes);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        // 4. 64 x 32 x kc on the FP64 tensor cores: 4 warps, warp tile 32 x 16
        double acc[4][2][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        const int ksteps = (ft.kc + 3) >> 2;
        for (int ks = 0; ks < ksteps; ++ks) {
            const int k4 = ks * 4;
            double af[4], bf[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = As[(k4 + ac) * kFlowLDA + wm0 + i * 8 + ar];
#pragma unroll
            for (int j = 0; j < 2; ++j) bf[j] = Bs[(wn0 + j * 8 + ar) * kFlowLDB + k4 + ac];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        // 5. accumulate into the destination and publish
        const double sgn = *bad ? __longlong_as_double(0x7ff8000000000000LL) : (ft.kind == 0 ? 1.0 : -1.0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = wm0 + i * 8 + ar;
            if (row >= ft.rows) continue;
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int col = wn0 + j * 8 + 2 * ac + q;
                    if (col < ft.cols) atomicAdd(Cp + ft.cstart + row + (i64)col * ft.ldc, sgn * acc[i][j][q]);
                }
        }
        __threadfence();
        __syncthreads();   // also: nobody still reads As / Bs when the next tile's prefetch overwrites them
        if (tid == 0) atomicAdd(ctr + (i64)ft.sig_ctr * cs, 1);
    }
}

int32_t build_tri_flow_plan(bbm_handle_t h, bbm_sky_desc_t d, int upper, i64 ldb, i64 ldx, i64 nrhs) {
    bbm_sky_desc::Tri& t = d->tri[upper];
    const i64 fbk = bbm_opt(h, "trsm.flow_bk", 64);
    const i64 kFlowBK = fbk >= 128 ? 128 : (fbk <= 32 ? 32 : 64);   // k-slice per tile: 64; 32 = twice the tiles, each half as long; 128 falls back to the launch chain
    if (t.flow_ldb == ldb && t.flow_nrhs == nrhs && t.flow_ldx == ldx && t.flow_bk == kFlowBK && t.flow_ctr) return BBM_OK;
    const i64 N = d->Nr;
    std::vector<FlowTile> tiles;
    // Completion counters per 64-row tile of every block: cB[K][c] counts the update tiles that have landed in rows
    // [64c, 64c+64) of B_K, cX[K][i] the partial products of rows [64i, 64i+64) of X_K.  A tile waits only for the one
    // 64-row slice it reads (its k-slice), not for the whole block: no tile waits for the slowest of several hundred.
    std::vector<i64> toff(N + 1, 0);
    for (i64 K = 0; K < N; ++K) toff[K + 1] = toff[K] + (d->r[K] + kFlowBM - 1) / kFlowBM;
    const i64 nctr = toff[N];
    if (2 * nctr + 1 >= (i64)INT32_MAX) { t.flow_ok = false; t.flow_ldb = ldb; t.flow_nrhs = nrhs; t.flow_ldx = ldx; t.flow_bk = kFlowBK; return BBM_OK; }
    std::vector<int> nA(nctr, 0), nB(nctr, 0);   // targets: cX / cB
    bool even = !(ldb & 1) && !(ldx & 1) && kFlowBK <= kFlowBM;   // (a k-slice must lie inside one 64-row slice: one counter to wait for)
    for (i64 s = 0; s < N; ++s) {
        const i64 K = upper ? N - 1 - s : s;
        const i64 rK = d->r[K];
        if (rK == 0) continue;
        const i64 wo = t.woff[K], wl = t.wld[K];
        for (i64 i0 = 0; i0 < rK; i0 += kFlowBM) {
            const i64 nr = std::min<i64>(kFlowBM, rK - i0);
            const i64 klo = upper ? i0 : 0, khi = upper ? rK : std::min<i64>(rK, i0 + kFlowBM);   // skip the zero triangle of the inverse
            for (i64 k0 = klo; k0 < khi; k0 += kFlowBK) {
                const i64 kc = std::min<i64>(kFlowBK, khi - k0);
                for (i64 j0 = 0; j0 < nrhs; j0 += kFlowBN) {
                    FlowTile ft{wo + i0 + k0 * wl, wl, d->R[K] + k0 + j0 * ldb, ldb, d->R[K] + i0 + j0 * ldx, ldx,
                                (int32_t)nr, (int32_t)std::min<i64>(kFlowBN, nrhs - j0), (int32_t)kc, 0,
                                (int32_t)(toff[K] + k0 / kFlowBM), 0, (int32_t)(nctr + toff[K] + i0 / kFlowBM), 0};
                    even = even && !(ft.astart & 1) && !(ft.lda & 1) && !(ft.bstart & 1);
                    tiles.push_back(ft);
                    ++nA[toff[K] + i0 / kFlowBM];
                }
            }
        }
        // update tiles, nearest block row first (it is the one the next step waits for)
        const i64 Kfirst = upper ? K - 1 : K + 1, Klast = upper ? d->klo[K] : d->khi[K];
        for (i64 Kr = Kfirst; upper ? Kr >= Klast : Kr <= Klast; Kr += upper ? -1 : 1) {
            const i64 rr = d->r[Kr];
            if (rr == 0) continue;
            const i64 bs = sky_start(d, Kr, K);
            for (i64 i0 = 0; i0 < rr; i0 += kFlowBM) {
                const i64 nr = std::min<i64>(kFlowBM, rr - i0);
                for (i64 k0 = 0; k0 < rK; k0 += kFlowBK) {
                    const i64 kc = std::min<i64>(kFlowBK, rK - k0);
                    for (i64 j0 = 0; j0 < nrhs; j0 += kFlowBN) {
                        FlowTile ft{bs + i0 + k0 * d->S[K], d->S[K], d->R[K] + k0 + j0 * ldx, ldx, d->R[Kr] + i0 + j0 * ldb, ldb,
                                    (int32_t)nr, (int32_t)std::min<i64>(kFlowBN, nrhs - j0), (int32_t)kc, 1,
                                    (int32_t)(nctr + toff[K] + k0 / kFlowBM), 0, (int32_t)(toff[Kr] + i0 / kFlowBM), 0};
                        even = even && !(ft.astart & 1) && !(ft.lda & 1) && !(ft.bstart & 1);
                        tiles.push_back(ft);
                        ++nB[toff[Kr] + i0 / kFlowBM];
                    }
                }
            }
        }
    }
    for (FlowTile& ft : tiles) ft.wait_val = ft.kind == 0 ? nB[ft.wait_ctr] : nA[ft.wait_ctr - nctr];
    if (t.flow_tiles) { cudaFree(t.flow_tiles); t.flow_tiles = nullptr; }
    if (t.flow_ctr) { cudaFree(t.flow_ctr); t.flow_ctr = nullptr; }
    t.flow_ok = even && !tiles.empty() && tiles.size() < (size_t)INT32_MAX;
    t.flow_ntiles = (i64)tiles.size();
    if (t.flow_ok) {
        FlowTile* dt = nullptr;
        int32_t rc = upload(h, tiles, &dt);
        if (rc != BBM_OK) return rc;
        t.flow_tiles = dt;
    }
    t.flow_nctr = 2 * nctr;
    cudaError_t e = cudaMalloc(&t.flow_ctr, sizeof(int) * (size_t)(2 * nctr + 1) * kFlowCtrStride);
    if (e != cudaSuccess) { cudaGetLastError(); return bbm_fail(h, BBM_ERR_ALLOC, "cudaMalloc(counters): %s", cudaGetErrorString(e)); }
    t.flow_ldb = ldb; t.flow_nrhs = nrhs; t.flow_ldx = ldx; t.flow_bk = kFlowBK;
    return BBM_OK;
}

// ||T_KK||_inf and ||inv(T_KK)||_inf per diagonal block (one CTA each): the condition guard of the inverse-based paths.
// blk holds 5 entries per block: offset of T_KK in `data`, its leading dimension, offset of the inverse in W, its
// leading dimension, rows.
__global__ void __launch_bounds__(256) sky_tri_cond_kernel(const double* __restrict__ data, const double* __restrict__ W, const i64* __restrict__ blk,
                                                           int upper, int unit, double* __restrict__ out) {
    const i64 K = blockIdx.x;
    const i64 ds = blk[5 * K], ld = blk[5 * K + 1], wo = blk[5 * K + 2], wld = blk[5 * K + 3];
    const int r = (int)blk[5 * K + 4];
    __shared__ double red[2][256];
    double mt = 0.0, mw = 0.0;
    for (int i = threadIdx.x; i < r; i += 256) {
        double st = 0.0, sw = 0.0;
        const int jlo = upper ? i : 0, jhi = upper ? r : i + 1;
        for (int j = jlo; j < jhi; ++j) {
            const double tv = (unit && j == i) ? 1.0 : data[ds + i + (i64)j * ld];
            st += fabs(tv);
            sw += fabs(W[wo + i + (i64)j * wld]);
        }
        mt = fmax(mt, st); mw = fmax(mw, sw);
    }
    red[0][threadIdx.x] = mt; red[1][threadIdx.x] = mw;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            red[0][threadIdx.x] = fmax(red[0][threadIdx.x], red[0][threadIdx.x + o]);
            red[1][threadIdx.x] = fmax(red[1][threadIdx.x], red[1][threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[2 * K] = red[0][0]; out[2 * K + 1] = red[1][0]; }
}

int32_t tri_condition(bbm_handle_t h, bbm_sky_desc_t d, int upper, int unit, const double* d_data, const double* W, double* cond_max) {
    bbm_sky_desc::Tri& t = d->tri[upper];
    const i64 N = d->Nr;
    *cond_max = 0.0;
    if (N == 0) return BBM_OK;
    if (!t.d_cond) {
        std::vector<i64> blk(5 * N);
        for (i64 K = 0; K < N; ++K) {
            blk[5 * K] = d->r[K] ? sky_start(d, K, K) : 0; blk[5 * K + 1] = d->S[K];
            blk[5 * K + 2] = t.woff[K]; blk[5 * K + 3] = t.wld[K]; blk[5 * K + 4] = d->r[K];
        }
        // one allocation: [2N doubles of results | 5N table entries]
        cudaError_t e = cudaMalloc(&t.d_cond, sizeof(double) * (size_t)(2 * N) + sizeof(i64) * (size_t)(5 * N));
        if (e != cudaSuccess) { cudaGetLastError(); return bbm_fail(h, BBM_ERR_ALLOC, "cudaMalloc(condition table): %s", cudaGetErrorString(e)); }
        BBM_CUDA(h, cudaMemcpy(t.d_cond + 2 * N, blk.data(), sizeof(i64) * (size_t)(5 * N), cudaMemcpyHostToDevice));
    }
    sky_tri_cond_kernel<<<(unsigned)N, 256, 0, h->stream>>>(d_data, W, reinterpret_cast<const i64*>(t.d_cond + 2 * N), upper, unit, t.d_cond);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "sky_tri_cond_kernel launch: %s", cudaGetErrorString(e));
    h->launches += 1;
    std::vector<double> out(2 * N);
    BBM_CUDA(h, cudaMemcpyAsync(out.data(), t.d_cond, sizeof(double) * (size_t)(2 * N), cudaMemcpyDeviceToHost, h->stream));
    BBM_CUDA(h, cudaStreamSynchronize(h->stream));
    double c = 0.0;
    for (i64 K = 0; K < N; ++K) {
        const double v = out[2 * K] * out[2 * K + 1];
        if (!(v == v)) { c = INFINITY; break; }   // NaN (a zero on a diagonal): not for the inverse path
        c = std::max(c, v);
    }
    *cond_max = c;
    return BBM_OK;
}

// One cooperative launch of the data-flow kernel over a tile list: static operands of kind-0 / kind-1 tiles from A0 / A1,
// dynamic operands and destinations Bm and X (kind 0: A0 * Bm -> X, +; kind 1: A1 * X -> Bm, -).
int32_t launch_flow(bbm_handle_t h, const void* flow_tiles, i64 flow_ntiles, int* flow_ctr, i64 flow_nctr, int bk,
                    const double* A1, const double* A0, double* Bm, double* X) {
    const void* func = bk == 32 ? reinterpret_cast<const void*>(sky_trsm_flow_kernel<32>) : reinterpret_cast<const void*>(sky_trsm_flow_kernel<64>);
    const size_t smem = bk == 32 ? flow_smem_bytes<32>() : flow_smem_bytes<64>();
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, func, 128, smem);
    if (e != cudaSuccess || occ < 1) { cudaGetLastError(); return bbm_fail(h, BBM_ERR_CUDA, "data-flow solve: occupancy query failed"); }
    const i64 cap = bbm_opt(h, "trsm.flow_ctas_per_sm", 0);
    if (cap > 0) occ = (int)std::min<i64>(occ, cap);
    const i64 grid = std::min<i64>(flow_ntiles, (i64)occ * h->num_sms);
    int cs = (int)std::min<i64>(kFlowCtrStride, std::max<i64>(1, bbm_opt(h, "trsm.flow_ctr_stride", kFlowCtrStride)));
    BBM_CUDA(h, cudaMemsetAsync(flow_ctr, 0, sizeof(int) * (size_t)(flow_nctr + 1) * cs, h->stream));
    const FlowTile* tiles = static_cast<const FlowTile*>(flow_tiles);
    i64 ntiles = flow_ntiles;
    int* ctr = flow_ctr;
    int err_idx = (int)flow_nctr;
    int sleep_ns = (int)std::max<i64>(0, bbm_opt(h, "trsm.flow_sleep_ns", 32));
    void* args[] = {(void*)&tiles, (void*)&ntiles, (void*)&A1, (void*)&A0, (void*)&Bm, (void*)&X, (void*)&ctr, (void*)&err_idx, (void*)&sleep_ns, (void*)&cs};
    // cooperative launch: the runtime guarantees that all CTAs are co-resident (the tiles wait for one another)
    e = cudaLaunchCooperativeKernel(func, dim3((unsigned)grid), dim3(128), args, smem, h->stream);
    if (e != cudaSuccess) { cudaGetLastError(); return bbm_fail(h, BBM_ERR_CUDA, "sky_trsm_flow_kernel launch: %s", cudaGetErrorString(e)); }
    h->launches += 1;
    return BBM_OK;
}

int32_t launch_trsm_flow(bbm_handle_t h, bbm_sky_desc_t d, int upper, const double* d_data, const double* W, double* d_B, double* X) {
    bbm_sky_desc::Tri& t = d->tri[upper];
    return launch_flow(h, t.flow_tiles, t.flow_ntiles, t.flow_ctr, t.flow_nctr, (int)t.flow_bk, d_data, W, d_B, X);
}

// ------------------------------------------------------------------------------------------------
// Transposed matvec  Y <- alpha * A' * X + beta * Y  (mul!(y, A', x), src/adjtransblockbanded.jl:2-7).
// A column of a panel is one contiguous run of S_J entries and its x operand the contiguous rows the panel
// spans, so y[c] is a plain dot product of two unit-stride streams: one warp per column, 8 columns per CTA,
// the matrix read exactly once, no atomics.  Summation order: lanes stride the rows, then a shuffle tree.
// ------------------------------------------------------------------------------------------------
template <typename T, int NR>
__global__ void __launch_bounds__(256) sky_gemv_t_kernel(const T* __restrict__ data, const i64* __restrict__ items, T alpha, const T* __restrict__ X,
                                                        i64 ldx, int q0, T beta, T* __restrict__ Y, i64 ldy) {
    const i64* it = items + (i64)blockIdx.x * 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= (int)it[4]) return;
    const i64 S = it[1];
    const T* col = data + it[0] + (i64)warp * S;
    const T* x = X + it[2] + (i64)q0 * ldx;
    T acc[NR];
#pragma unroll
    for (int q = 0; q < NR; ++q) acc[q] = T(0);
    for (i64 i = lane; i < S; i += 32) {
        const T a = col[i];
#pragma unroll
        for (int q = 0; q < NR; ++q) acc[q] = fma(a, x[i + (i64)q * ldx], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < NR; ++q) {
        T v = acc[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
            T* y = Y + it[3] + warp + (i64)(q0 + q) * ldy;
            *y = beta == T(0) ? alpha * v : fma(beta, *y, alpha * v);
        }
    }
}

template <typename T>
int32_t sky_mul_t_impl(bbm_handle_t h, bbm_sky_desc_t d, T alpha, const T* d_data, const T* d_X, i64 ldx, i64 nrhs, T beta, T* d_Y, i64 ldy) {
    BBM_REQUIRE_HANDLE(h);
    if (!sky_valid(d) || d->h != h) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BlockSkylineMatrix descriptor");
    if (nrhs < 0) return bbm_fail(h, BBM_ERR_ARG, "nrhs < 0");
    if (nrhs > 0 && (ldx < std::max<i64>(1, d->m) || ldy < std::max<i64>(1, d->n)))
        return bbm_fail(h, BBM_ERR_DIM, "DimensionMismatch: A' is %lld x %lld, ldx=%lld, ldy=%lld", (long long)d->n, (long long)d->m, (long long)ldx, (long long)ldy);
    if (d->n == 0 || nrhs == 0) return BBM_OK;
    if (!d_X && d->m > 0) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    if (!d_Y || (!d_data && d->numel > 0)) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    BBM_CUDA(h, cudaSetDevice(h->device));
    if (!d->d_titems) {
        std::vector<i64> items;
        for (i64 J = 0; J < d->Nc; ++J) {
            // empty panels (no stored block, or only empty block rows) still own their y entries: S = 0 gives beta*y
            const i64 S = d->S[J];
            const i64 xrow = S > 0 ? d->R[d->klo[J]] : 0;
            for (i64 j0 = 0; j0 < d->c[J]; j0 += 8)
                items.insert(items.end(), {d->off[J] + j0 * S, S, xrow, d->C[J] + j0, std::min<i64>(8, d->c[J] - j0)});
        }
        d->ntitems = (i64)items.size() / 5;
        int32_t rc = upload(h, items, &d->d_titems);
        if (rc != BBM_OK) return rc;
    }
    if (d->ntitems == 0) return BBM_OK;
    for (i64 q0 = 0; q0 < nrhs;) {
        const int nr = nrhs - q0 >= 4 ? 4 : 1;
        if (nr == 4) sky_gemv_t_kernel<T, 4><<<(unsigned)d->ntitems, 256, 0, h->stream>>>(d_data, d->d_titems, alpha, d_X, ldx, (int)q0, beta, d_Y, ldy);
        else sky_gemv_t_kernel<T, 1><<<(unsigned)d->ntitems, 256, 0, h->stream>>>(d_data, d->d_titems, alpha, d_X, ldx, (int)q0, beta, d_Y, ldy);
        h->launches += 1;
        q0 += nr;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "sky_gemv_t_kernel launch: %s", cudaGetErrorString(e));
    return BBM_OK;
}

// ------------------------------------------------------------------------------------------------
// QR solve: lmul!(F.Q', B) and ldiv!(F, B) for F = qr(A) of a block-banded matrix
// (src/blockskylineqr.jl:77-145).  F.factors is block-skyline with bandwidths (l, l+u): block column K
// holds R in and above its diagonal block and, in block rows K..K+l, the Householder vectors of panel K
// (LAPACK packing, unit diagonal implied); F.tau their scalar factors.
//
// The reference applies panel after panel with ormqr (one reflector after the other inside).  Here each
// panel's product of reflectors is put into compact-WY form once, for all panels in parallel,
//      Q_K = I - V T V',      Q_K' B = B - V (T' V') B = B - V (YT B),      YT = T' V'
// with T from the Gram matrix G = V'V: 32x32 diagonal blocks by the larft recurrence
// T[0:j,j] = -tau_j T[0:j,0:j] G[0:j,j], then doubling merges T12 = -T11 G12 T22 as batched tensor-core
// products (stored transposed: T'21 = -T'22 (G21 T'11)).  What stays sequential is two split-k DMMA
// products per panel, chained with programmatic dependent launches like the triangular solve:
//      W_K = YT_K B[KR,:]   (c_K x nrhs),       B[KR,:] -= V_K W_K.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sky_qr_extract_kernel(const double* __restrict__ fac, const i64* __restrict__ pan, double* __restrict__ ws) {
    const i64* q = pan + (i64)blockIdx.y * 6;
    const i64 fstart = q[0], ld = q[1], S = q[2], c = q[3], vc = q[4], vt = q[5];
    for (i64 idx = (i64)blockIdx.x * 256 + threadIdx.x; idx < S * c; idx += (i64)gridDim.x * 256) {
        const i64 i = idx % S, j = idx / S;
        const double v = i < j ? 0.0 : (i == j ? 1.0 : fac[fstart + i + j * ld]);
        ws[vc + i + j * S] = v;   // V, unit lower trapezoidal, zeros above
        ws[vt + j + i * c] = v;   // V'
    }
}

// one warp per block of up to 32 reflectors: T by the larft recurrence, written transposed
__global__ void __launch_bounds__(64) sky_qr_t32_kernel(const double* __restrict__ tau, const i64* __restrict__ blk, i64 nblk, double* __restrict__ ws) {
    __shared__ double Ts[2][32][33], Gs[2][32][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const i64 b = (i64)blockIdx.x * 2 + warp;
    if (b >= nblk) return;
    const i64 g0 = blk[b * 5], ld = blk[b * 5 + 1], t0 = blk[b * 5 + 2], tau0 = blk[b * 5 + 3];
    const int sz = (int)blk[b * 5 + 4];
    for (int j = 0; j < sz; ++j) {
        Gs[warp][lane][j] = lane < sz ? ws[g0 + lane + (i64)j * ld] : 0.0;
        Ts[warp][lane][j] = 0.0;
    }
    __syncwarp();
    for (int j = 0; j < sz; ++j) {
        const double tj = tau[tau0 + j];
        double sum = 0.0;
        if (lane < j)
            for (int k = lane; k < j; ++k) sum = fma(Ts[warp][lane][k], Gs[warp][k][j], sum);
        __syncwarp();
        if (lane < j) Ts[warp][lane][j] = -tj * sum;
        if (lane == j) Ts[warp][j][j] = tj;
        __syncwarp();
    }
    for (int i = 0; i < sz; ++i)   // T'[j,i] = T[i,j]; lane = j
        if (lane < sz && lane >= i) ws[t0 + lane + (i64)i * ld] = Ts[warp][i][lane];
}

int32_t build_qr_plan(bbm_handle_t h, bbm_sky_desc_t d) {
    bbm_sky_desc::Qr& q = d->qr;
    if (q.built) return BBM_OK;
    const i64 L = d->Nc > 0 ? d->l[0] : 0;
    for (i64 J = 0; J < d->Nc; ++J)
        if (d->l[J] != L) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "QR solve needs a uniform lower block bandwidth (BlockBandedMatrix factors)");
    if (d->Nr < d->Nc) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "wide QR (fewer block rows than block columns) is not taken by the device path");
    q.ncols = d->Nc;
    const i64 P = q.ncols;
    q.srows.assign(P, 0); q.vc.assign(P, 0); q.vt.assign(P, 0); q.g.assign(P, 0); q.tt.assign(P, 0); q.yt.assign(P, 0); q.xx.assign(P, 0);
    auto even = [](i64 v) { return (v + 1) & ~i64(1); };
    i64 off = 0;
    std::vector<i64> pan((size_t)P * 6), blk;
    // layout: [V | V' | YT] per panel, then [G], then [T'] (one contiguous range, zeroed per call), then [X]
    for (i64 K = 0; K < P; ++K) {
        if (d->c[K] > 0 && !sky_has(d, K, K)) return bbm_fail(h, BBM_ERR_BAND, "BandError: the factors have no diagonal block (%lld,%lld)", (long long)K, (long long)K);
        const i64 K1 = std::min(K + L, d->Nr - 1);
        const i64 S = d->R[K1 + 1] - d->R[K], c = d->c[K];
        q.srows[K] = S;
        q.vc[K] = off; off = even(off + S * c);
        q.vt[K] = off; off = even(off + S * c);
        q.yt[K] = off; off = even(off + S * c);
        q.maxel = std::max(q.maxel, S * c);
        pan[K * 6 + 0] = c > 0 ? sky_start(d, K, K) : 0; pan[K * 6 + 1] = d->S[K]; pan[K * 6 + 2] = S; pan[K * 6 + 3] = c;
        pan[K * 6 + 4] = q.vc[K]; pan[K * 6 + 5] = q.vt[K];
    }
    for (i64 K = 0; K < P; ++K) { q.g[K] = off; off = even(off + d->c[K] * d->c[K]); }
    q.tt_lo = off;
    for (i64 K = 0; K < P; ++K) { q.tt[K] = off; off = even(off + d->c[K] * d->c[K]); }
    q.tt_hi = off;
    for (i64 K = 0; K < P; ++K) { q.xx[K] = off; off = even(off + d->c[K] * d->c[K]); }
    q.total = off;
    MmBatchHost gram, ytb;
    std::vector<MmBatchHost> lev;
    for (i64 K = 0; K < P; ++K) {
        const i64 c = d->c[K], S = q.srows[K];
        if (c == 0 || S == 0) continue;
        gram.add(q.g[K], c, c, c, MmSeg{q.vt[K], c, q.vc[K], S, S});                 // G = V' V
        ytb.add(q.yt[K], c, c, S, MmSeg{q.tt[K], c, q.vt[K], c, c});                 // YT = T' V'
        for (i64 i0 = 0; i0 < c; i0 += 32) {
            blk.push_back(q.g[K] + i0 + i0 * c); blk.push_back(c); blk.push_back(q.tt[K] + i0 + i0 * c);
            blk.push_back(d->C[K] + i0); blk.push_back(std::min<i64>(32, c - i0));
        }
        int lv = 0;
        for (i64 hh = 32; hh < c; hh *= 2, ++lv) {
            if ((size_t)(2 * lv + 2) > lev.size()) lev.resize(2 * lv + 2);
            for (i64 i0 = 0; i0 + hh < c; i0 += 2 * hh) {
                const i64 h1 = hh, h2 = std::min(hh, c - i0 - hh);
                const i64 lo = (i0 + hh) + i0 * c, dg = (i0 + hh) + (i0 + hh) * c;
                lev[2 * lv].add(q.xx[K] + lo, c, h2, h1, MmSeg{q.g[K] + lo, c, q.tt[K] + i0 + i0 * c, c, h1});     // X = G21 T'11
                lev[2 * lv + 1].add(q.tt[K] + lo, c, h2, h1, MmSeg{q.tt[K] + dg, c, q.xx[K] + lo, c, h2});        // T'21 = -T'22 X
            }
        }
    }
    q.nblk = (i64)blk.size() / 5;
    int32_t rc = upload(h, pan, &q.d_pan);
    if (rc == BBM_OK) rc = upload(h, blk, &q.d_blk);
    if (rc == BBM_OK) rc = upload_batch(h, gram, &q.gram);
    if (rc == BBM_OK) rc = upload_batch(h, ytb, &q.ytb);
    q.levels.resize(lev.size());
    for (size_t i = 0; i < lev.size() && rc == BBM_OK; ++i) rc = upload_batch(h, lev[i], &q.levels[i]);
    if (rc != BBM_OK) return rc;
    q.built = true;
    return BBM_OK;
}

int32_t build_qr_step_plan(bbm_handle_t h, bbm_sky_desc_t d, i64 ldb, i64 ldw, i64 nrhs) {
    bbm_sky_desc::Qr& q = d->qr;
    if (q.step_ldb == ldb && q.step_nrhs == nrhs && q.step_ldw == ldw && q.stepW.tiles) return BBM_OK;
    MmBatchHost w, b;
    const i64 P = q.ncols, ksplit = 64;
    q.ptrW.assign(P + 1, 0); q.ptrB.assign(P + 1, 0);
    for (i64 K = 0; K < P; ++K) {
        q.ptrW[K] = (int)w.tiles.size();
        q.ptrB[K] = (int)b.tiles.size();
        const i64 c = d->c[K], S = q.srows[K];
        if (c == 0 || S == 0) continue;
        for (i64 k0 = 0; k0 < S; k0 += ksplit)   // W_K += YT_K[:, k-slice] B[R_K + k-slice, :]
            w.add(d->C[K], ldw, c, nrhs, MmSeg{q.yt[K] + k0 * c, c, d->R[K] + k0, ldb, std::min(ksplit, S - k0)});
        for (i64 k0 = 0; k0 < c; k0 += ksplit)   // B[KR, :] -= V_K[:, k-slice] W_K[k-slice, :]
            b.add(d->R[K], ldb, S, nrhs, MmSeg{q.vc[K] + k0 * S, S, d->C[K] + k0, ldw, std::min(ksplit, c - k0)});
    }
    q.ptrW[P] = (int)w.tiles.size();
    q.ptrB[P] = (int)b.tiles.size();
    int32_t rc = upload_batch(h, w, &q.stepW);
    if (rc == BBM_OK) rc = upload_batch(h, b, &q.stepB);
    if (rc != BBM_OK) return rc;
    q.step_ldb = ldb; q.step_nrhs = nrhs; q.step_ldw = ldw;
    return BBM_OK;
}

// The sweep of lmul!(Q', B) as a tile list for the data-flow kernel (see sky_trsm_flow_kernel): per panel K
//     W_K        += YT_K[i-tile, k-slice] * B[R_K + k-slice, j-tile]      (kind 0; waits until that 64-row slice of B has
//                                                                          received every update of the panels before K)
//     B[R_K + i] -= V_K[i-tile, k-slice]  * W_K[k-slice, j-tile]          (kind 1; waits until ALL of W_K is complete: the
//                                                                          panel's rows of B are both read by its W tiles
//                                                                          and written by its update tiles)
// A 64-row slice of B is read by up to l+1 panels, each after a different number of updates: the counters are cumulative and
// every tile carries its own threshold.  Needs block sizes that are multiples of 64 (slices = tiles) and even strides.
int32_t build_qr_flow_plan(bbm_handle_t h, bbm_sky_desc_t d, i64 ldb, i64 ldw, i64 nrhs) {
    bbm_sky_desc::Qr& q = d->qr;
    if (q.flow_ldb == ldb && q.flow_nrhs == nrhs && q.flow_ldw == ldw && q.flow_ctr) return BBM_OK;
    if (q.flow_tiles) { cudaFree(q.flow_tiles); q.flow_tiles = nullptr; }
    if (q.flow_ctr) { cudaFree(q.flow_ctr); q.flow_ctr = nullptr; }
    q.flow_ldb = ldb; q.flow_nrhs = nrhs; q.flow_ldw = ldw; q.flow_ok = false;
    bool ok = !(ldb & 1) && !(ldw & 1);
    for (i64 K = 0; K < d->Nr && ok; ++K) ok = d->r[K] % kFlowBM == 0;
    for (i64 J = 0; J < d->Nc && ok; ++J) ok = d->c[J] % kFlowBM == 0;
    const i64 P = q.ncols;
    const i64 nrow = d->m / kFlowBM;                            // counters: 64-row slices of B, then one per panel (W_K complete)
    if (!ok || nrow + P + 1 >= (i64)INT32_MAX) {
        cudaError_t e0 = cudaMalloc(&q.flow_ctr, sizeof(int));   // (marks the plan as built for these strides)
        if (e0 != cudaSuccess) { cudaGetLastError(); return bbm_fail(h, BBM_ERR_ALLOC, "cudaMalloc(counters): %s", cudaGetErrorString(e0)); }
        return BBM_OK;
    }
    std::vector<FlowTile> tiles;
    std::vector<int> landed(nrow, 0);                           // update tiles that target a slice of B so far
    const i64 njt = (nrhs + kFlowBN - 1) / kFlowBN;
    for (i64 K = 0; K < P; ++K) {
        const i64 c = d->c[K], S = q.srows[K];
        if (c == 0 || S == 0) continue;
        const i64 r0 = d->R[K] / kFlowBM;
        for (i64 i0 = 0; i0 < c; i0 += kFlowBM)
            for (i64 k0 = 0; k0 < S; k0 += kFlowBM)
                for (i64 j0 = 0; j0 < nrhs; j0 += kFlowBN)
                    tiles.push_back(FlowTile{q.yt[K] + i0 + k0 * c, c, d->R[K] + k0 + j0 * ldb, ldb, d->C[K] + i0 + j0 * ldw, ldw,
                                             (int32_t)kFlowBM, (int32_t)std::min<i64>(kFlowBN, nrhs - j0), (int32_t)kFlowBM, 0,
                                             (int32_t)(r0 + k0 / kFlowBM), landed[r0 + k0 / kFlowBM], (int32_t)(nrow + K), 0});
        const int wdone = (int)((c / kFlowBM) * (S / kFlowBM) * njt);   // all W tiles of the panel
        for (i64 i0 = 0; i0 < S; i0 += kFlowBM) {
            for (i64 k0 = 0; k0 < c; k0 += kFlowBM)
                for (i64 j0 = 0; j0 < nrhs; j0 += kFlowBN)
                    tiles.push_back(FlowTile{q.vc[K] + i0 + k0 * S, S, d->C[K] + k0 + j0 * ldw, ldw, d->R[K] + i0 + j0 * ldb, ldb,
                                             (int32_t)kFlowBM, (int32_t)std::min<i64>(kFlowBN, nrhs - j0), (int32_t)kFlowBM, 1,
                                             (int32_t)(nrow + K), wdone, (int32_t)(r0 + i0 / kFlowBM), 0});
            landed[r0 + i0 / kFlowBM] += (int)((c / kFlowBM) * njt);
        }
    }
    if (tiles.empty() || tiles.size() >= (size_t)INT32_MAX) {
        cudaError_t e0 = cudaMalloc(&q.flow_ctr, sizeof(int));
        if (e0 != cudaSuccess) { cudaGetLastError(); return bbm_fail(h, BBM_ERR_ALLOC, "cudaMalloc(counters): %s", cudaGetErrorString(e0)); }
        return BBM_OK;
    }
    FlowTile* dt = nullptr;
    int32_t rc = upload(h, tiles, &dt);
    if (rc != BBM_OK) return rc;
    q.flow_tiles = dt;
    q.flow_ntiles = (i64)tiles.size();
    q.flow_nctr = nrow + P;
    cudaError_t e = cudaMalloc(&q.flow_ctr, sizeof(int) * (size_t)(q.flow_nctr + 1) * kFlowCtrStride);
    if (e != cudaSuccess) { cudaGetLastError(); return bbm_fail(h, BBM_ERR_ALLOC, "cudaMalloc(counters): %s", cudaGetErrorString(e)); }
    q.flow_ok = true;
    return BBM_OK;
}

int32_t sky_qr_lmul_adj_impl(bbm_handle_t h, bbm_sky_desc_t d, const double* d_fac, const double* d_tau, double* d_B, i64 ldb, i64 nrhs) {
    BBM_REQUIRE_HANDLE(h);
    if (!sky_valid(d) || d->h != h) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BlockSkylineMatrix descriptor");
    if (nrhs < 0) return bbm_fail(h, BBM_ERR_ARG, "nrhs < 0");
    if (nrhs > 0 && ldb < std::max<i64>(1, d->m)) return bbm_fail(h, BBM_ERR_DIM, "DimensionMismatch: Q is %lld x %lld, ldb=%lld", (long long)d->m, (long long)d->m, (long long)ldb);
    if (d->m == 0 || d->n == 0 || nrhs == 0) return BBM_OK;
    if (!d_fac || !d_tau || !d_B) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    BBM_CUDA(h, cudaSetDevice(h->device));
    int32_t rc = build_qr_plan(h, d);
    if (rc != BBM_OK) return rc;
    bbm_sky_desc::Qr& q = d->qr;
    const i64 ldw = std::max<i64>(2, (d->n + 1) & ~i64(1));
    rc = build_qr_step_plan(h, d, ldb, ldw, nrhs);
    if (rc == BBM_OK) rc = bbm_ws_reserve(h, &h->ws_z, &h->ws_z_bytes, (size_t)q.total * sizeof(double));
    if (rc == BBM_OK) rc = bbm_ws_reserve(h, &h->ws_y, &h->ws_y_bytes, (size_t)ldw * nrhs * sizeof(double));
    if (rc != BBM_OK) return rc;
    double* WS = static_cast<double*>(h->ws_z);
    double* W = static_cast<double*>(h->ws_y);
    BBM_CUDA(h, cudaMemsetAsync(WS + q.tt_lo, 0, (size_t)(q.tt_hi - q.tt_lo) * sizeof(double), h->stream));
    BBM_CUDA(h, cudaMemsetAsync(W, 0, (size_t)ldw * nrhs * sizeof(double), h->stream));
    // compact-WY form of all panels, in parallel
    const unsigned gx = (unsigned)std::min<i64>(64, (q.maxel + 255) / 256);
    sky_qr_extract_kernel<<<dim3(std::max(1u, gx), (unsigned)q.ncols), 256, 0, h->stream>>>(d_fac, q.d_pan, WS);
    h->launches += 1;
    rc = launch_mm64(h, q.gram, 0, q.gram.ntiles, WS, WS, WS, 1.0, 0.0, false);
    if (rc != BBM_OK) return rc;
    if (q.nblk > 0) {
        sky_qr_t32_kernel<<<(unsigned)((q.nblk + 1) / 2), 64, 0, h->stream>>>(d_tau, q.d_blk, q.nblk, WS);
        h->launches += 1;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "QR compact-WY kernels: %s", cudaGetErrorString(e));
    for (size_t lv = 0; lv + 1 < q.levels.size(); lv += 2) {
        rc = launch_mm64(h, q.levels[lv], 0, q.levels[lv].ntiles, WS, WS, WS, 1.0, 0.0, false);
        if (rc == BBM_OK) rc = launch_mm64(h, q.levels[lv + 1], 0, q.levels[lv + 1].ntiles, WS, WS, WS, -1.0, 0.0, false);
        if (rc != BBM_OK) return rc;
    }
    rc = launch_mm64(h, q.ytb, 0, q.ytb.ntiles, WS, WS, WS, 1.0, 0.0, false);
    if (rc != BBM_OK) return rc;
    // the sweep over the panels: one persistent data-flow kernel ("qr.flow", default on; many right-hand sides, block
    // sizes multiples of 64), else two dependent launches per panel
    if (bbm_opt(h, "qr.flow", 1) != 0 && nrhs >= 8 && (reinterpret_cast<uintptr_t>(d_B) % 16 == 0)) {
        rc = build_qr_flow_plan(h, d, ldb, ldw, nrhs);
        if (rc != BBM_OK) return rc;
        if (q.flow_ok) return launch_flow(h, q.flow_tiles, q.flow_ntiles, q.flow_ctr, q.flow_nctr, 64, WS, WS, d_B, W);
    }
    const bool pdl = bbm_opt(h, "trsm.pdl", 1) != 0;
    for (i64 K = 0; K < q.ncols; ++K) {
        rc = launch_mm64(h, q.stepW, q.ptrW[K], q.ptrW[K + 1] - q.ptrW[K], WS, d_B, W, 1.0, 1.0, true, K < 2, pdl);
        if (rc == BBM_OK) rc = launch_mm64(h, q.stepB, q.ptrB[K], q.ptrB[K + 1] - q.ptrB[K], WS, W, d_B, -1.0, 1.0, true, K < 2, pdl);
        if (rc != BBM_OK) return rc;
    }
    return BBM_OK;
}

// ------------------------------------------------------------------------------------------------
// qr!(A) for a BlockBandedMatrix already widened to the factors' bandwidths (l, l+u)
// (_blockbanded_qr!, src/blockskylineqr.jl:29-43): for every block column K, Householder QR of the panel
// V = A[K..K+l, K] and Q_K' applied to the blocks to its right.  The reference does the panel with the
// unblocked algorithm and the rest with ormqr; here a panel is processed in sub-panels of nb <= 32 columns:
//   sky_qr_panel_kernel  one CTA holds the (rows x nb) slab in shared memory, runs the nb reflector steps there
//                        (same reflector convention as LinearAlgebra.reflector!: nu = copysign(|x|, x0), ...),
//                        writes R / v / tau back in place and leaves the slab's compact-WY pieces V and
//                        YT = T' V' in a workspace slot;
//   two batched DMMA launches apply it to everything to the right (rest of the panel + trailing blocks):
//                        W = YT C,   C -= V W.
// The tile tables of all steps are built once per descriptor.
// ------------------------------------------------------------------------------------------------
constexpr int kQrThreads = 1024;   // one warp per slab column during a reflector step

__device__ __forceinline__ double qr_block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = red[lane];   // kQrThreads / 32 == 32 partial sums
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

__global__ void __launch_bounds__(kQrThreads) sky_qr_panel_kernel(double* __restrict__ data, double* __restrict__ tau, const i64* __restrict__ steps,
                                                                 i64 step, double* __restrict__ ws, i64 slotV0, i64 slotV1, i64 slotY0, i64 slotY1,
                                                                 i64 slotW0, i64 slotW1, i64 wlen) {
    extern __shared__ __align__(16) unsigned char qr_smem[];
    const i64* st = steps + step * 6;
    const i64 a0 = st[0], ld = st[1];
    const int rows = (int)st[2], w = (int)st[3];
    const i64 tau0 = st[4];
    const int par = (int)st[5];
    const int lds = rows | 1;                       // odd leading dimension: rows of one column hit all banks
    double* As = reinterpret_cast<double*>(qr_smem);            // [w][lds]
    double* Ts = As + (size_t)w * lds;                           // [32][33]  T (upper)
    double* Gs = Ts + 32 * 33;                                   // [32][33]  V'V (upper)
    double* red = Gs + 32 * 33;                                  // [32]
    double* sc = red + 32;                                       // [0] xi1, [1] tau
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = warp; c < w; c += kQrThreads / 32)        // one warp per slab column: no index division
        for (int r = lane; r < rows; r += 32) As[r + c * lds] = data[a0 + r + (i64)c * ld];
    for (int idx = tid; idx < 32 * 33; idx += kQrThreads) { Ts[idx] = 0.0; Gs[idx] = 0.0; }
    {   // the split-k products of this step accumulate into a zeroed W
        double* W = ws + (par ? slotW1 : slotW0);
        for (i64 idx = tid; idx < wlen; idx += kQrThreads) W[idx] = 0.0;
    }
    __syncthreads();
    for (int jj = 0; jj < w; ++jj) {
        double* x = As + jj + jj * lds;
        const int n = rows - jj;
        if (n <= 1) {                                // real element type: nothing to annihilate in the last row
            if (tid == 0 && n == 1) tau[tau0 + jj] = 0.0;
            if (tid == 0) sc[2 + jj] = 0.0;
            continue;
        }
        double ss = 0.0;
        for (int r = tid; r < n; r += kQrThreads) ss = fma(x[r], x[r], ss);
        ss = qr_block_sum(ss, red);
        if (tid == 0) {
            const double normu = sqrt(ss);
            double tj = 0.0, xi1 = 1.0;
            if (normu != 0.0) {
                xi1 = x[0];
                const double nu = copysign(normu, xi1);
                xi1 += nu;
                x[0] = -nu;
                tj = xi1 / nu;
            }
            sc[0] = xi1; sc[1] = tj; sc[2 + jj] = tj;
            tau[tau0 + jj] = tj;
        }
        __syncthreads();
        const double xi1 = sc[0], tj = sc[1];
        if (tj != 0.0)
            for (int r = 1 + tid; r < n; r += kQrThreads) x[r] /= xi1;
        __syncthreads();
        // one warp per other column: to the right the reflector is applied (reflectorApply!), to the left the
        // dot product v_t'v_jj is kept for the T factor
        if (warp < w && warp != jj && tj != 0.0) {
            double* a = As + jj + warp * lds;
            double v0 = 0.0, v1 = 0.0;
            for (int r = 1 + lane; r < n; r += 64) {
                v0 = fma(x[r], a[r], v0);
                if (r + 32 < n) v1 = fma(x[r + 32], a[r + 32], v1);
            }
            double v = v0 + v1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            v += a[0];
            if (warp > jj) {
                v *= tj;
                __syncwarp();
                if (lane == 0) a[0] -= v;
                for (int r = 1 + lane; r < n; r += 32) a[r] = fma(-x[r], v, a[r]);
            } else if (lane == 0) {
                Gs[warp * 33 + jj] = v;
            }
        }
        __syncthreads();
    }
    // R and the Householder vectors go back in place
    for (int c = warp; c < w; c += kQrThreads / 32)
        for (int r = lane; r < rows; r += 32) data[a0 + r + (i64)c * ld] = As[r + c * lds];
    __syncthreads();
    // clean V in shared memory (zeros above, ones on the diagonal); T by the larft recurrence; V and YT = T'V' out
    for (int c = warp; c < w; c += kQrThreads / 32)
        for (int r = lane; r <= c; r += 32) As[r + c * lds] = r == c ? 1.0 : 0.0;
    if (warp == 0) {
        for (int j = 0; j < w; ++j) {
            const double tj = sc[2 + j];
            double sum = 0.0;
            if (lane < j)
                for (int k = lane; k < j; ++k) sum = fma(Ts[lane * 33 + k], Gs[k * 33 + j], sum);
            __syncwarp();
            if (lane < j) Ts[lane * 33 + j] = -tj * sum;
            if (lane == j) Ts[j * 33 + j] = tj;
            __syncwarp();
        }
    }
    __syncthreads();
    double* Vout = ws + (par ? slotV1 : slotV0);   // rows x w, ld = rows
    double* Yout = ws + (par ? slotY1 : slotY0);   // w x rows, ld = w
    for (int c = warp; c < w; c += kQrThreads / 32)
        for (int r = lane; r < rows; r += 32) Vout[r + (i64)c * rows] = As[r + c * lds];
    for (int r = warp; r < rows; r += kQrThreads / 32) {       // YT[i, r] = sum_{k <= i} T[k, i] V[r, k]; lane = i, a warp writes one contiguous column
        if (lane < w) {
            double v = 0.0;
            for (int k = 0; k <= lane; ++k) v = fma(Ts[k * 33 + lane], As[r + k * lds], v);
            Yout[lane + (i64)r * w] = v;
        }
    }
}

// Helpers of the register-resident panel step.  A slab column lives in the registers of one warp (lane owns rows lane,
// lane+32, ...); a slab has at most 32 columns, so row jj is element 0 of lane jj and only element 0 can lie above it.
//
// The reflector's scalars sit on the critical path of every step (all lanes compute them redundantly): square root,
// quotient and reciprocal come from the hardware's 20-bit seeds with two Newton steps and one residual correction each
// (the quotient-from-reciprocal sequence of the IEEE routines without their special-case branches: ~12 dependent
// instructions instead of ~35).  Arguments outside [1e-280, 1e280] take the library routines.
__device__ __forceinline__ double qr_rcp(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ double qr_sqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a * y, y, 1.0);
    y = fma(0.5 * y, e, y);
    e = fma(-a * y, y, 1.0);
    y = fma(0.5 * y, e, y);
    const double sq = a * y;
    return fma(fma(-sq, sq, a), 0.5 * y, sq);
}
// One Householder reflector on a slab column, convention of LinearAlgebra.reflector!: nu = copysign(|x|, x_jj),
// x_jj <- -nu, x[jj+1:] scaled by 1 / (x_jj + nu) (a multiplication by the reciprocal, as LAPACK's dlarfg scales),
// tau = (x_jj + nu) / nu.  Publishes v (0 above row jj, 1 in row jj) in xb and tau in sc[jj] / tau_out[jj].
template <int NR>
__device__ __forceinline__ void qr_reflect(double (&col)[NR], int jj, int rows, int lane, double* __restrict__ xb, double* __restrict__ sc,
                                           double* __restrict__ tau_out) {
    double tj = 0.0;
    if (rows - jj > 1) {                 // real element type: nothing to annihilate in the last row
        double s0 = lane >= jj ? col[0] * col[0] : 0.0, s1 = col[1] * col[1], s2 = col[2] * col[2], s3 = col[3] * col[3];
#pragma unroll
        for (int i = 4; i < NR; i += 4) {
            s0 = fma(col[i], col[i], s0);
            s1 = fma(col[i + 1], col[i + 1], s1);
            s2 = fma(col[i + 2], col[i + 2], s2);
            s3 = fma(col[i + 3], col[i + 3], s3);
        }
        double ss = (s0 + s1) + (s2 + s3);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        double xi1 = __shfl_sync(0xffffffffu, col[0], jj);
        if (ss != 0.0) {
            const bool fast = ss > 1e-280 && ss < 1e280;
            const double normu = fast ? qr_sqrt(ss) : sqrt(ss);
            const double nu = copysign(normu, xi1);
            xi1 += nu;                   // same sign: |xi1| >= |nu| > 0
            double rinv;
            if (fast && fabs(xi1) < 1e280) {
                const double rnu = qr_rcp(nu);
                rinv = qr_rcp(xi1);
                const double q = xi1 * rnu;
                tj = fma(fma(-q, nu, xi1), rnu, q);
            } else {
                tj = xi1 / nu;
                rinv = 1.0 / xi1;
            }
            if (lane > jj) col[0] *= rinv;
            else if (lane == jj) col[0] = -nu;
#pragma unroll
            for (int i = 1; i < NR; ++i) col[i] *= rinv;
        }
    }
    xb[lane] = lane < jj ? 0.0 : lane == jj ? 1.0 : col[0];
#pragma unroll
    for (int i = 1; i < NR; ++i) xb[lane + 32 * i] = col[i];
    if (lane == 0) { sc[jj] = tj; tau_out[jj] = tj; }
}

// Register-resident variant of the panel step (the default when a slab has at most 1024 rows).  A warp keeps CPW slab
// columns in registers for the whole step (lane owns rows lane, lane+32, ...: NR values per column), so a reflector
// step is
//   owner warp : |x| by shuffles, nu / tau, x scaled, v published in shared memory (double-buffered)
//   one CTA barrier
//   every warp : v into registers once, then per column v'a (one pass over registers, one shuffle reduction) and
//                a -= tau (v'a) v (one more pass)
// -- 1 barrier, NR shared-memory loads and ~4 NR FP64 instructions per warp instead of 4 barriers and shared-memory
// traffic for every element.  The step is a chain of dependent FP64 instructions (ncu: ~13 cycles per issued
// instruction and warp, FP64 pipe 17 % busy), so what counts is the NUMBER of instructions on the owner -> barrier ->
// update -> next owner path.  History on a 512 x 32 slab: shared-memory slab kernel 140 us; one column per warp, v
// re-read for the update, spills: 122 us (LSU queue; profiles/r02_sky_qr_panel_reg_v1_ncu_details.txt); two columns
// per warp, v in registers: 62 us; a two-barrier variant that overlaps the owner's scalar chain with everybody's dot
// products on the raw column was SLOWER (70 us: more instructions per step than the overlap saved).
// Column jj-1 of T (larft recurrence) is computed during step jj by a warp that is half a panel away from the owner.
// The tail writes R / v / tau and V from the registers and forms YT = T'V' from V in shared memory: lane = slab row,
// a warp task = 32 rows x 16 columns of YT, staged through shared memory for contiguous stores.
template <int NR, int NWARP, int CPW>
__global__ void __launch_bounds__(NWARP * 32, 1) sky_qr_panel_reg_kernel(double* __restrict__ data, double* __restrict__ tau, const i64* __restrict__ steps,
                                                                       i64 step, double* __restrict__ ws, i64 slotV0, i64 slotV1, i64 slotY0, i64 slotY1,
                                                                       i64 slotW0, i64 slotW1, i64 wlen) {
    extern __shared__ __align__(16) unsigned char qr_smem[];
    constexpr int NT = NWARP * 32, RMAX = NR * 32, TW = 34, NB = NWARP * CPW, NH = NB / 16, TPW = NR * NH / NWARP, SB = NB + 1;
    static_assert(NR % 16 == 0 && (NB == 16 || NB == 32) && TPW >= 1 && (NR * NH) % NWARP == 0, "slab shape");
    const i64* st = steps + step * 6;
    const i64 a0 = st[0], ld = st[1];
    const int rows = (int)st[2], w = (int)st[3];
    const i64 tau0 = st[4];
    const int par = (int)st[5];
    const int lds = rows | 1;
    double* xs = reinterpret_cast<double*>(qr_smem);            // [2][RMAX] the current reflector (0 above, 1 on the diagonal)
    double* Ts = xs + 2 * RMAX;                                  // [32][TW]  T (upper)
    double* Gs = Ts + 32 * TW;                                   // [32][33]  V'V (upper)
    double* sc = Gs + 32 * 33;                                   // [32] tau | [32] 1 / (x_jj + nu), padded to 72
    double* As = sc + 72;                                        // [w][lds] V, later [rows][SB] YT'
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double a[CPW][NR];
#pragma unroll
    for (int s = 0; s < CPW; ++s) {
        const int c = warp * CPW + s;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const int r = lane + 32 * i;
            a[s][i] = (c < w && r < rows) ? data[a0 + r + (i64)c * ld] : 0.0;
        }
    }
    for (int idx = tid; idx < 32 * TW; idx += NT) Ts[idx] = 0.0;
    for (int idx = tid; idx < 32 * 33; idx += NT) Gs[idx] = 0.0;
    {   // the split-k products of this step accumulate into a zeroed W
        double* W = ws + (par ? slotW1 : slotW0);
        for (i64 idx = tid; idx < wlen; idx += NT) W[idx] = 0.0;
    }
    __syncthreads();
    auto t_column = [&](int j) {     // T[0:j, j] = -tau_j T[0:j, 0:j] G[0:j, j], T[j, j] = tau_j; lane = row of T
        const double tj = sc[j];
        double s0 = 0.0, s1 = 0.0;
        if (lane < j) {
            int k = lane;
            for (; k + 1 < j; k += 2) {
                s0 = fma(Ts[lane * TW + k], Gs[k * 33 + j], s0);
                s1 = fma(Ts[lane * TW + k + 1], Gs[(k + 1) * 33 + j], s1);
            }
            if (k < j) s0 = fma(Ts[lane * TW + k], Gs[k * 33 + j], s0);
        }
        __syncwarp();
        if (lane < j) Ts[lane * TW + j] = -tj * (s0 + s1);
        if (lane == j) Ts[j * TW + j] = tj;
    };
    for (int jj = 0; jj < w; ++jj) {
        double* xb = xs + (jj & 1) * RMAX;
        if (warp == jj / CPW) {                      // (explicit branches: the column array must stay in registers)
            if (CPW == 1 || jj % CPW == 0) qr_reflect<NR>(a[0], jj, rows, lane, xb, sc, tau + tau0);
            else qr_reflect<NR>(a[CPW - 1], jj, rows, lane, xb, sc, tau + tau0);
        }
        __syncthreads();
        const double tj = sc[jj];
        if (tj != 0.0 && warp * CPW < w) {
            constexpr bool kKeepX = (CPW + 1) * NR * 2 <= 100;   // v stays in registers between the two passes if it fits
            double xv[kKeepX ? NR : 1];
            if constexpr (kKeepX) {
#pragma unroll
                for (int i = 0; i < NR; ++i) xv[i] = xb[lane + 32 * i];
            }
            double v[CPW];
#pragma unroll
            for (int s = 0; s < CPW; ++s) {
                double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
#pragma unroll
                for (int i = 0; i < NR; i += 4) {
                    if constexpr (kKeepX) {
                        v0 = fma(xv[i], a[s][i], v0);
                        v1 = fma(xv[i + 1], a[s][i + 1], v1);
                        v2 = fma(xv[i + 2], a[s][i + 2], v2);
                        v3 = fma(xv[i + 3], a[s][i + 3], v3);
                    } else {
                        v0 = fma(xb[lane + 32 * i], a[s][i], v0);
                        v1 = fma(xb[lane + 32 * (i + 1)], a[s][i + 1], v1);
                        v2 = fma(xb[lane + 32 * (i + 2)], a[s][i + 2], v2);
                        v3 = fma(xb[lane + 32 * (i + 3)], a[s][i + 3], v3);
                    }
                }
                v[s] = (v0 + v1) + (v2 + v3);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int s = 0; s < CPW; ++s) v[s] += __shfl_xor_sync(0xffffffffu, v[s], o);
            }
#pragma unroll
            for (int s = 0; s < CPW; ++s) {
                const int c = warp * CPW + s;
                if (c >= w || c == jj) continue;
                if (c > jj) {                        // reflectorApply! on the columns to the right
                    const double vt = v[s] * tj;         // (v is zero above row jj: those entries keep their value)
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        if constexpr (kKeepX) a[s][i] = fma(-xv[i], vt, a[s][i]);
                        else a[s][i] = fma(-xb[lane + 32 * i], vt, a[s][i]);
                    }
                } else if (lane == 0) {              // to the left: v_t'v_jj for the T factor
                    Gs[c * 33 + jj] = v[s];
                }
            }
        }
        if (jj > 0 && warp == ((jj / CPW) ^ (NWARP / 2))) t_column(jj - 1);
    }
    __syncthreads();
    if (warp == 0) t_column(w - 1);
    // R and the Householder vectors go back in place; V (zeros above, ones on the diagonal) to the workspace and to
    // shared memory
    double* Vout = ws + (par ? slotV1 : slotV0);   // rows x w, ld = rows
    double* Yout = ws + (par ? slotY1 : slotY0);   // w x rows, ld = w
#pragma unroll
    for (int s = 0; s < CPW; ++s) {
        const int c = warp * CPW + s;
        if (c >= w) continue;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const int r = lane + 32 * i;
            if (r < rows) {
                data[a0 + r + (i64)c * ld] = a[s][i];
                const double vc = r < c ? 0.0 : r == c ? 1.0 : a[s][i];
                Vout[r + (i64)c * rows] = vc;
                As[r + c * lds] = vc;
            }
        }
    }
    __syncthreads();
    // YT[i, r] = sum_{k <= i} T[k, i] V[r, k]: lane = r inside a group of 32 rows, a task = (row group, 16 values of i)
    double acc[TPW][16];
#pragma unroll
    for (int tt = 0; tt < TPW; ++tt) {
        const int task = warp + tt * NWARP;
        const int hh = NH == 2 ? (task >> 2) & 1 : 0;
        const int g = NH == 2 ? (task & 3) | ((task >> 3) << 2) : task;
        const int r = g * 32 + lane;
#pragma unroll
        for (int ii = 0; ii < 16; ++ii) acc[tt][ii] = 0.0;
        if (hh == 0) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if (k < w) {
                    const double v = r < rows ? As[r + k * lds] : 0.0;
#pragma unroll
                    for (int i2 = (k & ~1); i2 < 16; i2 += 2) {
                        const double2 t2 = *reinterpret_cast<const double2*>(Ts + k * TW + i2);
                        if (i2 >= k) acc[tt][i2] = fma(t2.x, v, acc[tt][i2]);
                        acc[tt][i2 + 1] = fma(t2.y, v, acc[tt][i2 + 1]);
                    }
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                if (k < w) {
                    const double v = r < rows ? As[r + k * lds] : 0.0;
#pragma unroll
                    for (int i2 = (k < 16 ? 16 : (k & ~1)); i2 < 32; i2 += 2) {
                        const double2 t2 = *reinterpret_cast<const double2*>(Ts + k * TW + i2);
                        if (i2 >= k) acc[tt][i2 - 16] = fma(t2.x, v, acc[tt][i2 - 16]);
                        acc[tt][i2 - 15] = fma(t2.y, v, acc[tt][i2 - 15]);
                    }
                }
            }
        }
    }
    __syncthreads();                                 // every V value has been read: the buffer now takes YT' ([row][i])
#pragma unroll
    for (int tt = 0; tt < TPW; ++tt) {
        const int task = warp + tt * NWARP;
        const int hh = NH == 2 ? (task >> 2) & 1 : 0;
        const int g = NH == 2 ? (task & 3) | ((task >> 3) << 2) : task;
        const int r = g * 32 + lane;
        if (r < rows) {
#pragma unroll
            for (int ii = 0; ii < 16; ++ii) As[r * SB + hh * 16 + ii] = acc[tt][ii];
        }
    }
    __syncthreads();
    for (int r = warp; r < rows; r += NWARP)
        if (lane < w) Yout[lane + (i64)r * w] = As[r * SB + lane];
}

int32_t build_qr_fact_plan(bbm_handle_t h, bbm_sky_desc_t d) {
    bbm_sky_desc::QrFact& q = d->qrf;
    if (q.built) return BBM_OK;
    const i64 L = d->Nc > 0 ? d->l[0] : 0, U = d->Nc > 0 ? d->u[0] : 0;
    for (i64 J = 0; J < d->Nc; ++J)
        if (d->l[J] != L || d->u[J] != U) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "qr! needs uniform block bandwidths (BlockBandedMatrix)");
    const i64 P = std::min(d->Nr, d->Nc);
    i64 maxS = 0, maxcols = 0;
    for (i64 K = 0; K < P; ++K) {
        if (d->c[K] > 0 && d->r[K] > 0 && !sky_has(d, K, K)) return bbm_fail(h, BBM_ERR_BAND, "BandError: no diagonal block (%lld,%lld)", (long long)K, (long long)K);
        const i64 K1 = std::min(K + L, d->Nr - 1);
        maxS = std::max(maxS, d->R[K1 + 1] - d->R[K]);
        i64 cols = 0;
        for (i64 J = K; J <= std::min(K + U, d->Nc - 1); ++J) cols += d->c[J];
        maxcols = std::max(maxcols, cols);
    }
    // variant 1 / 2: the register-resident slab kernel, 16 warps x 2 columns x 16 rows per lane / 16 warps x 1 column x 32 rows per lane;
    // 0: the shared-memory slab kernel (taller panels, or qr.panel = 1)
    q.variant = bbm_opt(h, "qr.panel", 0) == 1 ? 0 : maxS <= 512 ? 1 : maxS <= 1024 ? 2 : 0;
    q.nb = q.variant == 1 ? 32 : q.variant == 2 ? 16 : maxS <= 704 ? 32 : maxS <= 1472 ? 16 : 8;
    if (q.variant == 0 && (size_t)q.nb * (maxS | 1) * 8 + 2 * 32 * 33 * 8 + 1024 > std::min<size_t>(h->smem_optin, 220 * 1024))
        return bbm_fail(h, BBM_ERR_UNSUPPORTED, "qr!: panels of %lld rows exceed the shared-memory slab", (long long)maxS);
    q.maxrows = maxS;
    auto even = [](i64 v) { return (v + 1) & ~i64(1); };
    i64 off = 0;
    for (int p = 0; p < 2; ++p) {
        q.slotV[p] = off; off = even(off + maxS * q.nb);
        q.slotY[p] = off; off = even(off + maxS * q.nb);
        q.slotW[p] = off; off = even(off + (i64)q.nb * maxcols);
    }
    q.wlen = (i64)q.nb * maxcols;
    q.ws_total = off;
    std::vector<i64> steps;
    MmBatchHost bw, bc;
    q.ptrW.clear(); q.ptrC.clear(); q.midW.clear(); q.midC.clear();
    struct Slab { i64 K, j0, w; };
    std::vector<Slab> slabs;
    for (i64 K = 0; K < P; ++K) {
        const i64 K1 = std::min(K + L, d->Nr - 1);
        const i64 S = d->R[K1 + 1] - d->R[K], c = d->c[K];
        if (S == 0 || c == 0) continue;
        for (i64 j0 = 0; j0 < std::min(S, c); j0 += q.nb) slabs.push_back(Slab{K, j0, std::min<i64>(q.nb, std::min(S, c) - j0)});
    }
    i64 nsteps = 0;
    for (size_t si = 0; si < slabs.size(); ++si) {
        const i64 K = slabs[si].K, j0 = slabs[si].j0, w = slabs[si].w;
        const i64 K1 = std::min(K + L, d->Nr - 1);
        const i64 S = d->R[K1 + 1] - d->R[K], c = d->c[K];
        const i64 vstart = sky_start(d, K, K), ld = d->S[K];
        const i64 rows = S - j0;
        // the columns the next step factorises: block column Kn, [j0n, j0n + wn) -- always a prefix of one piece below
        const i64 Kn = si + 1 < slabs.size() ? slabs[si + 1].K : -1, wn = si + 1 < slabs.size() ? slabs[si + 1].w : 0;
        const int par = (int)(nsteps & 1);
        steps.insert(steps.end(), {vstart + j0 + j0 * ld, ld, rows, w, d->C[K] + j0, (i64)par});
        // tiles of this step: first the part on the next slab's columns ([ptr[t], mid[t])), then the rest ([mid[t], ptr[t+1]))
        for (int pass = 0; pass < 2; ++pass) {
            (pass == 0 ? q.ptrW : q.midW).push_back((int)bw.tiles.size());
            (pass == 0 ? q.ptrC : q.midC).push_back((int)bc.tiles.size());
            i64 woff = q.slotW[par];
            auto part = [&](bool emit, i64 cstart, i64 ldc, i64 cols) {   // rows x cols target with top-left at data[cstart]
                if (cols <= 0) return;
                if (emit) {
                    for (i64 k0 = 0; k0 < rows; k0 += 64)                                       // W += YT[:, k-slice] C[k-slice, :]
                        bw.add(woff, w, w, cols, MmSeg{q.slotY[par] + k0 * w, w, cstart + k0, ldc, std::min<i64>(64, rows - k0)});
                    bc.add(cstart, ldc, rows, cols, MmSeg{q.slotV[par], rows, woff, w, w});   // C -= V W
                }
                woff += w * cols;
            };
            auto piece = [&](i64 J, i64 cstart, i64 ldc, i64 cols) {
                const i64 ncols = J == Kn ? std::min(wn, cols) : 0;
                part(pass == 0, cstart, ldc, ncols);
                part(pass == 1, cstart + ncols * ldc, ldc, cols - ncols);
            };
            piece(K, vstart + j0 + (j0 + w) * ld, ld, c - j0 - w);                          // rest of the panel
            for (i64 J = K + 1; J <= std::min(K + U, d->Nc - 1); ++J) {
                if (d->c[J] == 0) continue;
                if (!sky_has(d, K, J) || !sky_has(d, K1, J)) return bbm_fail(h, BBM_ERR_BAND, "BandError: the factors structure must have upper block bandwidth l+u");
                piece(J, sky_start(d, K, J) + j0, d->S[J], d->c[J]);
            }
        }
        ++nsteps;
    }
    q.ptrW.push_back((int)bw.tiles.size());
    q.ptrC.push_back((int)bc.tiles.size());
    q.nsteps = nsteps;
    int32_t rc = upload(h, steps, &q.d_steps);
    if (rc == BBM_OK) rc = upload_batch(h, bw, &q.bW);
    if (rc == BBM_OK) rc = upload_batch(h, bc, &q.bC);
    if (rc != BBM_OK) return rc;
    q.built = true;
    return BBM_OK;
}

int32_t sky_qr_fact_impl(bbm_handle_t h, bbm_sky_desc_t d, double* d_data, double* d_tau) {
    BBM_REQUIRE_HANDLE(h);
    if (!sky_valid(d) || d->h != h) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BlockSkylineMatrix descriptor");
    const i64 ntau = d->Nr < d->Nc ? d->m : d->n;
    if (ntau == 0) return BBM_OK;
    if (!d_data || !d_tau) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    BBM_CUDA(h, cudaSetDevice(h->device));
    int32_t rc = build_qr_fact_plan(h, d);
    if (rc != BBM_OK) return rc;
    bbm_sky_desc::QrFact& q = d->qrf;
    rc = bbm_ws_reserve(h, &h->ws_w, &h->ws_w_bytes, (size_t)q.ws_total * sizeof(double));
    if (rc != BBM_OK) return rc;
    double* WS = static_cast<double*>(h->ws_w);
    BBM_CUDA(h, cudaMemsetAsync(d_tau, 0, (size_t)ntau * sizeof(double), h->stream));
    size_t smem = ((size_t)q.nb * (q.maxrows | 1) + 2 * 32 * 33 + 32 + 2 + 32 + 2) * sizeof(double);
    if (q.variant != 0) {
        const size_t rmax = q.variant == 1 ? 512 : 1024;
        smem = (2 * rmax + 32 * 34 + 32 * 33 + 72 + std::max<size_t>((size_t)q.nb * (q.maxrows | 1), (size_t)q.maxrows * (q.nb + 1))) * sizeof(double);
    }
    auto kern = q.variant == 1 ? sky_qr_panel_reg_kernel<16, 16, 2> : q.variant == 2 ? sky_qr_panel_reg_kernel<32, 16, 1> : sky_qr_panel_kernel;
    BBM_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool pdl = bbm_opt(h, "trsm.pdl", 1) != 0;
    const unsigned nthreads = q.variant != 0 ? 512 : kQrThreads;
    // Look-ahead ("qr.lookahead" = 1; default OFF): only the columns the NEXT step factorises are updated on the main
    // stream, in front of that step; the rest of the update (two launches over all the trailing blocks) runs on a second
    // stream beside the next panel step, which occupies one SM.
    //   main  : P(t) ........ wait R(t-1) . Wn(t) Cn(t) . P(t+1) ...
    //   side  :      wait P(t) . Wr(t) Cr(t) -> R(t)
    // R(t-1) updated the columns Wn(t)/Cn(t) touch, hence the wait; P(t+2) reuses the workspace slots R(t) reads and
    // is ordered behind it through the same wait one step later.  Measured (512 blocks of 256, (1,1)): 345 ms with,
    // 336 ms without: a dependent launch costs ~10 us whether it carries 4 tiles or 90, so the two small launches left
    // on the critical path cost what the two big ones did, and the events add host work.  Kept for wider trailing
    // updates (large l + u), where the rest of the update is no longer launch-bound.
    const bool look = bbm_opt(h, "qr.lookahead", 0) != 0;
    if (look) {
        if (!h->side_stream) {
            int lo = 0, hi = 0;
            BBM_CUDA(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
            BBM_CUDA(h, cudaStreamCreateWithPriority(&h->side_stream, cudaStreamNonBlocking, lo));
            for (int i = 0; i < 2; ++i) BBM_CUDA(h, cudaEventCreateWithFlags(&h->side_ev[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < 4; ++i)
            if (!h->qr_ev[i]) BBM_CUDA(h, cudaEventCreateWithFlags(&h->qr_ev[i], cudaEventDisableTiming));
    }
    for (i64 t = 0; t < q.nsteps; ++t) {
        kern<<<1, nthreads, smem, h->stream>>>(d_data, d_tau, q.d_steps, t, WS, q.slotV[0], q.slotV[1], q.slotY[0], q.slotY[1], q.slotW[0], q.slotW[1], q.wlen);
        h->launches += 1;
        if (look) {
            BBM_CUDA(h, cudaEventRecord(h->qr_ev[t & 1], h->stream));
            if (t > 0) BBM_CUDA(h, cudaStreamWaitEvent(h->stream, h->qr_ev[2 + ((t - 1) & 1)], 0));
        }
        if (!look) {                                 // one W launch and one C launch over all the tiles of the step
            rc = launch_mm64(h, q.bW, q.ptrW[t], q.ptrW[t + 1] - q.ptrW[t], WS, d_data, WS, 1.0, 1.0, true, t < 2, pdl);
            if (rc == BBM_OK) rc = launch_mm64(h, q.bC, q.ptrC[t], q.ptrC[t + 1] - q.ptrC[t], WS, WS, d_data, -1.0, 1.0, false, t < 2, pdl);
            if (rc != BBM_OK) return rc;
            continue;
        }
        rc = launch_mm64(h, q.bW, q.ptrW[t], q.midW[t] - q.ptrW[t], WS, d_data, WS, 1.0, 1.0, true, t < 2, pdl);
        if (rc == BBM_OK) rc = launch_mm64(h, q.bC, q.ptrC[t], q.midC[t] - q.ptrC[t], WS, WS, d_data, -1.0, 1.0, false, t < 2, pdl);
        cudaStream_t rs = h->side_stream;
        BBM_CUDA(h, cudaStreamWaitEvent(rs, h->qr_ev[t & 1], 0));
        if (rc == BBM_OK) rc = launch_mm64(h, q.bW, q.midW[t], q.ptrW[t + 1] - q.midW[t], WS, d_data, WS, 1.0, 1.0, true, t < 2, pdl, rs);
        if (rc == BBM_OK) rc = launch_mm64(h, q.bC, q.midC[t], q.ptrC[t + 1] - q.midC[t], WS, WS, d_data, -1.0, 1.0, false, t < 2, pdl, rs);
        BBM_CUDA(h, cudaEventRecord(h->qr_ev[2 + (t & 1)], rs));
        if (rc != BBM_OK) return rc;
    }
    if (look && q.nsteps > 0) BBM_CUDA(h, cudaStreamWaitEvent(h->stream, h->qr_ev[2 + ((q.nsteps - 1) & 1)], 0));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "sky_qr_panel_kernel: %s", cudaGetErrorString(e));
    return BBM_OK;
}

template <typename T>
int32_t sky_trsm_impl(bbm_handle_t h, bbm_sky_desc_t d, int32_t uplo, int32_t unit, const T* d_data, T* d_B, i64 ldb, i64 nrhs);

// ------------------------------------------------------------------------------------------------
// Y <- alpha * X * A + beta * Y for a dense column-major X (p x m) and a block-skyline A (`X*A`, `MulAdd(X, A)`,
// test/test_linalg.jl:56-57; [upstream] the block loop Y[:, J] += X[:, K] A_KJ over K in blockcolsupport(A, J)).
// The in-band blocks of block column J are ONE contiguous column-major panel, so the whole column is a single dense
// product Y[:, C_J : C_J+c_J) = X[:, R_klo : R_khi+1) * panel_J: the panel is the k-contiguous operand the tile kernels
// want for "B", X the row-contiguous one for "A".  One batched launch (FP64: DMMA; FP32: the register-tiled SIMT kernel).
// ------------------------------------------------------------------------------------------------
template <typename T>
int32_t sky_rmul_impl(bbm_handle_t h, bbm_sky_desc_t d, T alpha, const T* d_X, i64 ldx, i64 p, const T* d_data, T beta, T* d_Y, i64 ldy) {
    BBM_REQUIRE_HANDLE(h);
    if (!sky_valid(d) || d->h != h) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BlockSkylineMatrix descriptor");
    if (p < 0) return bbm_fail(h, BBM_ERR_ARG, "negative row count");
    if (p > 0 && (ldx < p || ldy < p)) return bbm_fail(h, BBM_ERR_DIM, "DimensionMismatch: X is %lld x %lld, ldx=%lld, ldy=%lld", (long long)p, (long long)d->m, (long long)ldx, (long long)ldy);
    if (p == 0 || d->n == 0) return BBM_OK;
    if (!d_Y || (d->m > 0 && !d_X) || (d->numel > 0 && !d_data)) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    if ((const void*)d_Y == (const void*)d_X) return bbm_fail(h, BBM_ERR_ARG, "mul!: destination must not alias an operand");
    BBM_CUDA(h, cudaSetDevice(h->device));
    bbm_sky_desc::Rmul& r = d->rmul;
    if (r.p != p || r.ldx != ldx || r.ldy != ldy || r.elem != (int)sizeof(T)) {
        MmBatchHost b;
        b.tile = sizeof(T) == 8 ? 64 : 128;
        bool a4 = true;
        for (i64 J = 0; J < d->Nc; ++J) {
            if (d->c[J] == 0) continue;
            const i64 S = d->S[J];
            if (S == 0) { b.add_empty(d->C[J] * ldy, ldy, p, d->c[J]); continue; }
            const MmSeg sg{d->R[d->klo[J]] * ldx, ldx, d->off[J], S, S};
            a4 = a4 && !(sg.astart & 3) && !(sg.lda & 3) && !(sg.bstart & 3) && !(sg.ldb & 3);
            b.add(d->C[J] * ldy, ldy, p, d->c[J], sg);
        }
        int32_t rc = upload_batch(h, b, &r.batch);
        if (rc != BBM_OK) { r.p = -1; return rc; }
        r.p = p; r.ldx = ldx; r.ldy = ldy; r.elem = (int)sizeof(T); r.aligned4 = a4;
    }
    if (r.batch.ntiles == 0) return BBM_OK;
    if constexpr (sizeof(T) == 8) {
        return launch_mm64(h, r.batch, 0, r.batch.ntiles, (const double*)d_X, (const double*)d_data, (double*)d_Y, (double)alpha, (double)beta, false);
    } else {
        MmParams<float> pf{(const float*)d_X, (const float*)d_data, (float*)d_Y, (const MmCBlk*)r.batch.cblks, (const MmSeg*)r.batch.segs,
                           (const MmTile*)r.batch.tiles, (float)alpha, (float)beta};
        const bool al = r.aligned4 && (reinterpret_cast<uintptr_t>(d_X) % 16 == 0) && (reinterpret_cast<uintptr_t>(d_data) % 16 == 0);
        cudaError_t e = cudaSuccess;
        auto go = [&](auto kern) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSgSmem);
            if (e == cudaSuccess) { kern<<<r.batch.ntiles, 256, kSgSmem, h->stream>>>(pf); e = cudaGetLastError(); }
        };
        if (al) go(sky_sgemm_simt_kernel<true>);
        else go(sky_sgemm_simt_kernel<false>);
        if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "X*A product launch: %s", cudaGetErrorString(e));
        h->launches += 1;
        return BBM_OK;
    }
}

// ------------------------------------------------------------------------------------------------
// QL (ql!, lmul!(F.Q', B), ldiv!(F::QL, B); src/blockskylineqr.jl:51-72, 96-111, 149-157) by reversal.
// LAPACK's geqlf/ormql are geqrf/ormqr on the matrix with rows and columns reversed, and for block-skyline
// storage that reversed matrix is simply the flat `data` vector read backwards (panels in reverse order, every
// panel reversed, block sizes reversed, l and u swapped).  So the QL calls flip the operands, run the QR
// kernels above on the reversed descriptor, and flip back; factors and tau come out in geqlf's packing.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sky_flip_kernel(double* __restrict__ a, i64 n) {
    const i64 i = (i64)blockIdx.x * 256 + threadIdx.x;
    if (i < n / 2) { const double t = a[i]; a[i] = a[n - 1 - i]; a[n - 1 - i] = t; }
}
__global__ void __launch_bounds__(256) sky_flip_copy_kernel(const double* __restrict__ src, double* __restrict__ dst, i64 n) {
    const i64 i = (i64)blockIdx.x * 256 + threadIdx.x;
    if (i < n) dst[i] = src[n - 1 - i];
}
__global__ void __launch_bounds__(256) sky_flip_rows_kernel(double* __restrict__ B, i64 m, i64 ldb) {
    const i64 i = (i64)blockIdx.x * 256 + threadIdx.x;
    double* b = B + (i64)blockIdx.y * ldb;
    if (i < m / 2) { const double t = b[i]; b[i] = b[m - 1 - i]; b[m - 1 - i] = t; }
}

int32_t sky_reversed_desc(bbm_handle_t h, bbm_sky_desc_t d, bbm_sky_desc_t* out) {
    if (d->Nr != d->Nc) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "block-banded QL on the device takes square block structures");
    if (!d->rev) {
        std::vector<i64> r(d->r.rbegin(), d->r.rend()), c(d->c.rbegin(), d->c.rend());
        std::vector<i64> l(d->u.rbegin(), d->u.rend()), u(d->l.rbegin(), d->l.rend());
        bbm_sky_desc_t rv = nullptr;
        int32_t rc = bbm_sky_desc_create(h, d->Nr, d->Nc, r.data(), c.data(), l.data(), u.data(), &rv);
        if (rc != BBM_OK) return rc;
        if (rv->numel != d->numel) { bbm_sky_desc_destroy(rv); return bbm_fail(h, BBM_ERR_UNSUPPORTED, "reversed structure has a different storage size"); }
        d->rev = rv;
    }
    *out = d->rev;
    return BBM_OK;
}

inline void sky_flip(bbm_handle_t h, double* a, i64 n) {
    if (n > 1) { sky_flip_kernel<<<(unsigned)((n / 2 + 255) / 256), 256, 0, h->stream>>>(a, n); h->launches += 1; }
}
inline void sky_flip_rows(bbm_handle_t h, double* B, i64 m, i64 ldb, i64 nrhs) {
    if (m > 1 && nrhs > 0) { sky_flip_rows_kernel<<<dim3((unsigned)((m / 2 + 255) / 256), (unsigned)nrhs), 256, 0, h->stream>>>(B, m, ldb); h->launches += 1; }
}

int32_t sky_ql_fact_impl(bbm_handle_t h, bbm_sky_desc_t d, double* d_data, double* d_tau) {
    BBM_REQUIRE_HANDLE(h);
    if (!sky_valid(d) || d->h != h) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BlockSkylineMatrix descriptor");
    if (d->n == 0) return BBM_OK;
    if (!d_data || !d_tau) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    BBM_CUDA(h, cudaSetDevice(h->device));
    bbm_sky_desc_t rv = nullptr;
    int32_t rc = sky_reversed_desc(h, d, &rv);
    if (rc != BBM_OK) return rc;
    sky_flip(h, d_data, d->numel);
    rc = sky_qr_fact_impl(h, rv, d_data, d_tau);
    sky_flip(h, d_data, d->numel);
    sky_flip(h, d_tau, d->n);
    return rc;
}

// Q' B for the packed Q of a QL; with solve != 0 the LowerTriangular solve follows (ldiv!(F::QL, B))
int32_t sky_ql_apply_impl(bbm_handle_t h, bbm_sky_desc_t d, const double* d_fac, const double* d_tau, double* d_B, i64 ldb, i64 nrhs, int solve) {
    BBM_REQUIRE_HANDLE(h);
    if (!sky_valid(d) || d->h != h) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BlockSkylineMatrix descriptor");
    if (nrhs < 0) return bbm_fail(h, BBM_ERR_ARG, "nrhs < 0");
    if (nrhs > 0 && ldb < std::max<i64>(1, d->m)) return bbm_fail(h, BBM_ERR_DIM, "DimensionMismatch: Q is %lld x %lld, ldb=%lld", (long long)d->m, (long long)d->m, (long long)ldb);
    if (d->m == 0 || d->n == 0 || nrhs == 0) return BBM_OK;
    if (!d_fac || !d_tau || !d_B) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    BBM_CUDA(h, cudaSetDevice(h->device));
    bbm_sky_desc_t rv = nullptr;
    int32_t rc = sky_reversed_desc(h, d, &rv);
    if (rc != BBM_OK) return rc;
    bool match = true;
    for (i64 K = 0; match && K < d->Nr; ++K) match = d->r[K] == d->c[K];
    if (!match) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "QL solve on the device needs matching diagonal blocks");
    // reversed copies of the factors and tau (the caller's stay untouched)
    const size_t fb = ((size_t)d->numel * sizeof(double) + 255) / 256 * 256;
    rc = bbm_ws_reserve(h, &h->ws_x, &h->ws_x_bytes, fb + (size_t)d->n * sizeof(double) + 256);
    if (rc != BBM_OK) return rc;
    double* frev = static_cast<double*>(h->ws_x);
    double* trev = reinterpret_cast<double*>(static_cast<char*>(h->ws_x) + fb);
    sky_flip_copy_kernel<<<(unsigned)((d->numel + 255) / 256), 256, 0, h->stream>>>(d_fac, frev, d->numel);
    sky_flip_copy_kernel<<<(unsigned)((d->n + 255) / 256), 256, 0, h->stream>>>(d_tau, trev, d->n);
    h->launches += 2;
    sky_flip_rows(h, d_B, d->m, ldb, nrhs);
    rc = sky_qr_lmul_adj_impl(h, rv, frev, trev, d_B, ldb, nrhs);
    sky_flip_rows(h, d_B, d->m, ldb, nrhs);
    if (rc != BBM_OK || !solve) return rc;
    return sky_trsm_impl<double>(h, d, 'L', 0, d_fac, d_B, ldb, nrhs);
}

template <typename T>
int32_t sky_trsm_impl(bbm_handle_t h, bbm_sky_desc_t d, int32_t uplo, int32_t unit, const T* d_data, T* d_B, i64 ldb, i64 nrhs) {
    BBM_REQUIRE_HANDLE(h);
    if (!sky_valid(d)) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BlockSkylineMatrix descriptor");
    if (uplo != 'U' && uplo != 'L' && uplo != 'u' && uplo != 'l') return bbm_fail(h, BBM_ERR_ARG, "uplo must be 'U' or 'L'");
    const int upper = (uplo == 'U' || uplo == 'u') ? 1 : 0;
    if (nrhs < 0) return bbm_fail(h, BBM_ERR_ARG, "nrhs < 0");
    // hasmatchingblocks (src/triblockbanded.jl:12,20): square diagonal blocks
    if (d->Nr != d->Nc || d->r != d->c)
        return bbm_fail(h, BBM_ERR_UNSUPPORTED, "triangular solve needs matching square diagonal blocks (the reference falls back to a generic solve)");
    if (nrhs > 0 && ldb < std::max<i64>(1, d->m)) return bbm_fail(h, BBM_ERR_DIM, "DimensionMismatch: T is %lld x %lld, ldb=%lld", (long long)d->m, (long long)d->m, (long long)ldb);
    for (i64 K = 0; K < d->Nr; ++K)
        if (d->r[K] > 0 && !sky_has(d, K, K)) return bbm_fail(h, BBM_ERR_BAND, "BandError: diagonal block %lld is outside the block band", (long long)K);
    if (d->m == 0 || nrhs == 0) return BBM_OK;
    if (!d_data || !d_B) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    if (nrhs > 8 * 65535) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "too many right-hand sides");
    BBM_CUDA(h, cudaSetDevice(h->device));
    int32_t rc = build_tri(h, d, upper);
    if (rc != BBM_OK) return rc;
    bbm_sky_desc::Tri& t = d->tri[upper];
    if (sizeof(T) == 8 && nrhs >= 16 && (bbm_opt(h, "trsm.kernel", 0) == 0 || bbm_opt(h, "trsm.kernel", 0) == 3)) {
        // the GEMM path stores an explicit inverse (and a level temporary) per diagonal block: take it only if
        // that workspace is a modest share of the free device memory, else solve by substitution
        double wbytes = 0;
        for (i64 K = 0; K < d->Nr; ++K) wbytes += 2.0 * 8.0 * (double)(d->r[K] + 1) * (double)(d->r[K] + 1);
        wbytes += 8.0 * (double)(d->m + 2) * (double)nrhs;
        size_t free_b = 0, total_b = 0;
        const size_t have = h->ws_z_bytes + h->ws_w_bytes + h->ws_y_bytes;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = 0; }
        if (wbytes <= 0.5 * ((double)free_b + (double)have)) {
            bool fallback = false;
            rc = sky_trsm_gemm(h, d, upper, unit ? 1 : 0, reinterpret_cast<const double*>(d_data), reinterpret_cast<double*>(d_B), ldb, nrhs, &fallback);
            if (rc != BBM_OK || !fallback) return rc;
            // ill-conditioned diagonal blocks: substitution below (backward stable like the reference's trsv)
        }
    }
    // pass 0: inverses of the 32x32 diagonal sub-blocks (workspace: 1024 entries per sub-block)
    rc = bbm_ws_reserve(h, &h->ws_x, &h->ws_x_bytes, (size_t)t.nsub * 1024 * sizeof(T));
    if (rc != BBM_OK) return rc;
    T* W = static_cast<T*>(h->ws_x);
    sky_tri_invdiag_kernel<T><<<(unsigned)((t.nsub + 3) / 4), 128, 0, h->stream>>>(d_data, t.d_sub, t.nsub, upper, unit ? 1 : 0, W);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "sky_tri_invdiag_kernel launch: %s", cudaGetErrorString(e));
    h->launches += 1;
    // right-hand sides per diag CTA: as many as keep b_K in shared memory, at most 8
    i64 maxr = 0;
    for (i64 K = 0; K < d->Nr; ++K) maxr = std::max(maxr, d->r[K]);
    const size_t budget = std::min<size_t>(h->smem_optin, 200 * 1024);
    int RC = nrhs > 128 ? 8 : 4;  // more CTAs along the right-hand sides shorten the sequential chain
    while (RC > 1 && (size_t)RC * maxr * sizeof(T) > budget) RC /= 2;
    if ((size_t)RC * maxr * sizeof(T) > budget) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "diagonal block of %lld rows does not fit the shared-memory solve", (long long)maxr);
    while (RC > 1 && nrhs < RC) RC /= 2;
    const size_t smem_max = (size_t)RC * maxr * sizeof(T);
    auto set_attr = [&](auto kern) { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max); };
    if (RC == 8) e = set_attr(sky_tri_diag_kernel<T, 8>);
    else if (RC == 4) e = set_attr(sky_tri_diag_kernel<T, 4>);
    else if (RC == 2) e = set_attr(sky_tri_diag_kernel<T, 2>);
    else e = set_attr(sky_tri_diag_kernel<T, 1>);
    if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    const unsigned gq = (unsigned)((nrhs + RC - 1) / RC);
    // L2 prefetch of the next block column's panel on a side stream, throttled to one step ahead
    const bool prefetch = bbm_opt(h, "trsm.prefetch", 1) != 0 && d->Nr > 2;
    if (prefetch && !h->side_stream) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        BBM_CUDA(h, cudaStreamCreateWithPriority(&h->side_stream, cudaStreamNonBlocking, lo));
        for (int i = 0; i < 2; ++i) BBM_CUDA(h, cudaEventCreateWithFlags(&h->side_ev[i], cudaEventDisableTiming));
    }
    if (prefetch) {  // the side stream starts after everything queued so far (d_data may still be being written)
        BBM_CUDA(h, cudaEventRecord(h->side_ev[0], h->stream));
        BBM_CUDA(h, cudaStreamWaitEvent(h->side_stream, h->side_ev[0], 0));
    }
    for (i64 s = 0; s < d->Nr; ++s) {
        const i64 K = upper ? d->Nr - 1 - s : s;
        const int rK = (int)d->r[K];
        if (rK == 0) continue;
        if (prefetch && s + 1 < d->Nr) {
            const i64 Kn = upper ? K - 1 : K + 1;
            const size_t bytes = (size_t)(d->off[Kn + 1] - d->off[Kn]) * sizeof(T);
            if (bytes > 0) {
                const unsigned blocks = (unsigned)std::min<size_t>(592, (bytes / 32 + 255) / 256 + 1);
                sky_l2_prefetch_kernel<<<blocks, 256, 0, h->side_stream>>>(reinterpret_cast<const char*>(d_data + d->off[Kn]), bytes, nullptr);
                h->launches += 1;
            }
        }
        const i64 dstart = sky_start(d, K, K);
        const T* Wk = W + t.sub_ptr[K] * 1024;
        T* Bk = d_B + d->R[K];
        {
            const size_t smem = (size_t)RC * rK * sizeof(T);
            if (RC == 8) sky_tri_diag_kernel<T, 8><<<gq, 256, smem, h->stream>>>(d_data, dstart, d->S[K], rK, upper, Wk, Bk, ldb, (int)nrhs);
            else if (RC == 4) sky_tri_diag_kernel<T, 4><<<gq, 256, smem, h->stream>>>(d_data, dstart, d->S[K], rK, upper, Wk, Bk, ldb, (int)nrhs);
            else if (RC == 2) sky_tri_diag_kernel<T, 2><<<gq, 256, smem, h->stream>>>(d_data, dstart, d->S[K], rK, upper, Wk, Bk, ldb, (int)nrhs);
            else sky_tri_diag_kernel<T, 1><<<gq, 256, smem, h->stream>>>(d_data, dstart, d->S[K], rK, upper, Wk, Bk, ldb, (int)nrhs);
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "sky_tri_diag_kernel launch: %s", cudaGetErrorString(e));
        h->launches += 1;
        const int ni = t.item_ptr[K + 1] - t.item_ptr[K];
        if (ni > 0) {
            GemvParams<T> p{d_data, d_B, d_B, ldb, ldb, t.d_items + t.item_ptr[K], t.d_ents, T(-1), T(1), (int)nrhs};
            // split the block's columns over CTAs until a few hundred are in flight
            const int groups = ni * (int)((nrhs + 15) / 16);
            int ksplit = 1;
            while (ksplit < 16 && groups * ksplit < 2 * h->num_sms && rK / (ksplit * 2) >= kGemvChunk) ksplit *= 2;
            if (nrhs <= 8) ksplit = 1;
            rc = launch_gemv<T>(h, p, ni, (int)nrhs, ksplit);
            if (rc != BBM_OK) return rc;
        }
        if (prefetch) {  // the prefetch of step s+2 may start once step s is done
            BBM_CUDA(h, cudaEventRecord(h->side_ev[s & 1], h->stream));
            BBM_CUDA(h, cudaStreamWaitEvent(h->side_stream, h->side_ev[s & 1], 0));
        }
    }
    if (prefetch) {  // join: nothing of this call is still in flight once the handle's stream is drained
        BBM_CUDA(h, cudaEventRecord(h->side_ev[0], h->side_stream));
        BBM_CUDA(h, cudaStreamWaitEvent(h->stream, h->side_ev[0], 0));
    }
    return BBM_OK;
}


// ------------------------------------------------------------------------------------------------
// Float32 QR / QL (src/blockskylineqr.jl:17-67 is generic in T): the factorisation and its solves run in Float64 on the
// FP64 tensor-core pipeline above -- the operands are widened into stream-ordered scratch buffers, the Float64 kernels run
// unchanged, and the results are rounded to Float32 ONCE at the end.  Every intermediate carries 29 more bits than the
// reference's Float32 arithmetic, so factors, tau and solutions agree with it to Float32 rounding (the parity tests apply
// the Float32 tolerance against a Float32 restatement of the reference's loops).
// ------------------------------------------------------------------------------------------------
template <typename S, typename D>
__global__ void __launch_bounds__(256) sky_convert_kernel(const S* __restrict__ src, i64 lds, D* __restrict__ dst, i64 ldd, i64 rows, i64 cols) {
    for (i64 c = blockIdx.y; c < cols; c += gridDim.y)
        for (i64 i = (i64)blockIdx.x * 256 + threadIdx.x; i < rows; i += (i64)gridDim.x * 256) dst[i + c * ldd] = static_cast<D>(src[i + c * lds]);
}
template <typename S, typename D>
void sky_convert(bbm_handle_t h, const S* src, i64 lds, D* dst, i64 ldd, i64 rows, i64 cols) {
    if (rows <= 0 || cols <= 0) return;
    const dim3 grid((unsigned)std::min<i64>((rows + 255) / 256, 4096), (unsigned)std::min<i64>(cols, 65535));
    sky_convert_kernel<S, D><<<grid, 256, 0, h->stream>>>(src, lds, dst, ldd, rows, cols);
    h->launches += 1;
}

struct ScratchF64 {   // stream-ordered Float64 scratch, released behind the work that uses it
    bbm_handle_t h;
    std::vector<void*> bufs;
    explicit ScratchF64(bbm_handle_t hh) : h(hh) {}
    double* get(i64 n) {
        void* p = nullptr;
        if (cudaMallocAsync(&p, sizeof(double) * (size_t)std::max<i64>(n, 1), h->stream) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        bufs.push_back(p);
        return static_cast<double*>(p);
    }
    ~ScratchF64() { for (void* p : bufs) cudaFreeAsync(p, h->stream); }
};

// which: 0 qr!, 1 ql!
int32_t sky_fact_f32(bbm_handle_t h, bbm_sky_desc_t d, float* d_data, float* d_tau, int which) {
    BBM_REQUIRE_HANDLE(h);
    if (!sky_valid(d) || d->h != h) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BlockSkylineMatrix descriptor");
    if (!d_data || !d_tau) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    BBM_CUDA(h, cudaSetDevice(h->device));
    const i64 ntau = d->Nr >= d->Nc ? d->n : d->m;
    ScratchF64 ws(h);
    double* D = ws.get(d->numel);
    double* Tau = ws.get(ntau);
    if (!D || !Tau) return bbm_fail(h, BBM_ERR_ALLOC, "Float64 scratch for the Float32 factorisation");
    sky_convert<float, double>(h, d_data, d->numel, D, d->numel, d->numel, 1);
    BBM_CUDA(h, cudaMemsetAsync(Tau, 0, sizeof(double) * (size_t)std::max<i64>(ntau, 1), h->stream));
    int32_t rc = which == 0 ? sky_qr_fact_impl(h, d, D, Tau) : sky_ql_fact_impl(h, d, D, Tau);
    if (rc != BBM_OK) return rc;
    sky_convert<double, float>(h, D, d->numel, d_data, d->numel, d->numel, 1);
    sky_convert<double, float>(h, Tau, ntau, d_tau, ntau, ntau, 1);
    return BBM_OK;
}

// which: 0 lmul!(Q', B) of a QR, 1 F \ B of a QR, 2 lmul!(Q', B) of a QL, 3 F \ B of a QL
int32_t sky_apply_f32(bbm_handle_t h, bbm_sky_desc_t d, const float* d_fac, const float* d_tau, float* d_B, i64 ldb, i64 nrhs, int which) {
    BBM_REQUIRE_HANDLE(h);
    if (!sky_valid(d) || d->h != h) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BlockSkylineMatrix descriptor");
    if (nrhs < 0) return bbm_fail(h, BBM_ERR_ARG, "nrhs < 0");
    if (nrhs > 0 && ldb < std::max<i64>(1, d->m)) return bbm_fail(h, BBM_ERR_DIM, "DimensionMismatch: Q is %lld x %lld, ldb=%lld", (long long)d->m, (long long)d->m, (long long)ldb);
    if (d->m == 0 || d->n == 0 || nrhs == 0) return BBM_OK;
    if (!d_fac || !d_tau || !d_B) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    BBM_CUDA(h, cudaSetDevice(h->device));
    const i64 ntau = d->Nr >= d->Nc ? d->n : d->m;
    const i64 ldd = std::max<i64>(2, (d->m + 1) & ~i64(1));
    ScratchF64 ws(h);
    double* D = ws.get(d->numel);
    double* Tau = ws.get(ntau);
    double* Bd = ws.get(ldd * nrhs);
    if (!D || !Tau || !Bd) return bbm_fail(h, BBM_ERR_ALLOC, "Float64 scratch for the Float32 solve");
    sky_convert<float, double>(h, d_fac, d->numel, D, d->numel, d->numel, 1);
    sky_convert<float, double>(h, d_tau, ntau, Tau, ntau, ntau, 1);
    sky_convert<float, double>(h, d_B, ldb, Bd, ldd, d->m, nrhs);
    int32_t rc;
    if (which == 0) rc = bbm_sky_qr_lmul_adj_f64(h, d, D, Tau, Bd, ldd, nrhs);
    else if (which == 1) rc = bbm_sky_qr_ldiv_f64(h, d, D, Tau, Bd, ldd, nrhs);
    else rc = sky_ql_apply_impl(h, d, D, Tau, Bd, ldd, nrhs, which == 3 ? 1 : 0);
    if (rc != BBM_OK) return rc;
    sky_convert<double, float>(h, Bd, ldd, d_B, ldb, d->m, nrhs);
    return BBM_OK;
}

}  // namespace

int32_t bbm_sky_mul_rows(bbm_handle_t h, bbm_sky_desc_t d, int elem, double alpha, const void* data, const void* X, i64 ldx, i64 nrhs, double beta,
                         void* Y, i64 ldy, i64 Klo, i64 Khi, int phase) {
    if (elem == 8) return sky_mul_impl<double>(h, d, alpha, (const double*)data, (const double*)X, ldx, nrhs, beta, (double*)Y, ldy, Klo, Khi, phase);
    return sky_mul_impl<float>(h, d, (float)alpha, (const float*)data, (const float*)X, ldx, nrhs, (float)beta, (float*)Y, ldy, Klo, Khi, phase);
}

int32_t bbm_sky_matmul_phase(bbm_handle_t h, bbm_sky_desc_t A, bbm_sky_desc_t B, bbm_sky_desc_t C, int elem, double alpha, const void* dA,
                             const void* dB, double beta, void* dC, int phase, i64 own_lo, i64 own_hi) {
    if (elem == 8) return sky_matmul_impl<double>(h, A, B, C, alpha, (const double*)dA, (const double*)dB, beta, (double*)dC, phase, own_lo, own_hi);
    return sky_matmul_impl<float>(h, A, B, C, (float)alpha, (const float*)dA, (const float*)dB, (float)beta, (float*)dC, phase, own_lo, own_hi);
}

bool bbm_sky_desc_view(bbm_sky_desc_t d, SkyDescView* out) {
    if (!sky_valid(d) || !out) return false;
    if (!d->d_struct) {
        std::vector<i64> t;
        t.insert(t.end(), d->off.begin(), d->off.end());
        t.insert(t.end(), d->S.begin(), d->S.end());
        t.insert(t.end(), d->klo.begin(), d->klo.end());
        t.insert(t.end(), d->khi.begin(), d->khi.end());
        t.insert(t.end(), d->R.begin(), d->R.end());
        t.insert(t.end(), d->C.begin(), d->C.end());
        if (cudaSetDevice(d->h->device) != cudaSuccess) return false;
        if (upload(d->h, t, &d->d_struct) != BBM_OK) return false;
    }
    const i64 Nc = d->Nc, Nr = d->Nr;
    SkyDescView v{};
    v.h = d->h; v.Nr = Nr; v.Nc = Nc; v.m = d->m; v.n = d->n; v.numel = d->numel;
    v.r = d->r.data(); v.c = d->c.data(); v.R = d->R.data(); v.C = d->C.data();
    v.klo = d->klo.data(); v.khi = d->khi.data(); v.off = d->off.data(); v.S = d->S.data();
    v.d_off = d->d_struct; v.d_S = v.d_off + (Nc + 1); v.d_klo = v.d_S + Nc; v.d_khi = v.d_klo + Nc;
    v.d_R = v.d_khi + Nc; v.d_C = v.d_R + (Nr + 1);
    *out = v;
    return true;
}

extern "C" {

int32_t bbm_sky_desc_create(bbm_handle_t h, int64_t Nr, int64_t Nc, const int64_t* r, const int64_t* c, const int64_t* lvec,
                            const int64_t* uvec, bbm_sky_desc_t* out) {
    BBM_REQUIRE_HANDLE(h);
    if (!out) return bbm_fail(h, BBM_ERR_ARG, "null out pointer");
    *out = nullptr;
    if (Nr < 0 || Nc < 0 || (Nr > 0 && !r) || (Nc > 0 && (!c || !lvec || !uvec))) return bbm_fail(h, BBM_ERR_ARG, "bad block axes");
    if (Nr > INT32_MAX / 4 || Nc > INT32_MAX / 4) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "too many blocks");
    bbm_sky_desc* d = new (std::nothrow) bbm_sky_desc();
    if (!d) return bbm_fail(h, BBM_ERR_ALLOC, "out of host memory");
    d->h = h; d->serial = g_serial.fetch_add(1); d->Nr = Nr; d->Nc = Nc;
    d->r.assign(r, r + Nr); d->c.assign(c, c + Nc);
    d->l.assign(lvec, lvec + Nc); d->u.assign(uvec, uvec + Nc);
    d->R.assign(Nr + 1, 0); d->C.assign(Nc + 1, 0);
    for (i64 K = 0; K < Nr; ++K) {
        if (r[K] < 0 || r[K] > INT32_MAX / 2) { delete d; return bbm_fail(h, BBM_ERR_ARG, "bad row block size"); }
        d->R[K + 1] = d->R[K] + r[K];
    }
    for (i64 J = 0; J < Nc; ++J) {
        if (c[J] < 0 || c[J] > INT32_MAX / 2) { delete d; return bbm_fail(h, BBM_ERR_ARG, "bad column block size"); }
        d->C[J + 1] = d->C[J] + c[J];
    }
    d->m = d->R[Nr]; d->n = d->C[Nc];
    // panels: bb_blockstarts / bb_blockstrides / bb_numentries (src/BlockSkylineMatrix.jl:6-43, 93-104)
    d->klo.assign(Nc, 0); d->khi.assign(Nc, -1); d->S.assign(Nc, 0); d->off.assign(Nc + 1, 0);
    for (i64 J = 0; J < Nc; ++J) {
        // KR = max(1,J-u[J]):min(J+l[J],N), possibly empty (src/BlockSkylineMatrix.jl:16, 35, 91)
        const i64 lo = std::max<i64>(0, J - uvec[J]), hi = std::min<i64>(Nr - 1, J + lvec[J]);
        d->klo[J] = lo;
        d->khi[J] = hi >= lo ? hi : lo - 1;
        d->S[J] = hi >= lo ? d->R[hi + 1] - d->R[lo] : 0;
        d->off[J + 1] = d->off[J] + d->S[J] * c[J];
        d->nblocks += std::max<i64>(0, d->khi[J] - d->klo[J] + 1);
    }
    d->numel = d->off[Nc];
    // matvec tables: per block row K the blocks (K,J) in ascending J; items = 128-row tiles
    std::vector<std::vector<i64>> rowJ(Nr);
    for (i64 J = 0; J < Nc; ++J)
        for (i64 K = d->klo[J]; K <= d->khi[J]; ++K)
            if (c[J] > 0) rowJ[K].push_back(J);
    std::vector<SkyEnt> ents;
    std::vector<SkyItem> items;
    d->row_item_ptr.assign(Nr + 1, 0);
    for (i64 K = 0; K < Nr; ++K) {
        d->row_item_ptr[K] = (int)items.size();
        if (r[K] == 0) continue;
        const int32_t e0 = (int32_t)ents.size();
        for (i64 J : rowJ[K]) ents.push_back(SkyEnt{sky_start(d, K, J), d->S[J], d->C[J], c[J]});
        const int32_t e1 = (int32_t)ents.size();
        for (i64 k0 = 0; k0 < r[K]; k0 += kGemvRows)
            items.push_back(SkyItem{d->R[K] + k0, k0, (int32_t)std::min<i64>(kGemvRows, r[K] - k0), e0, e1, 0});
    }
    d->nitems = (int)items.size();
    d->row_item_ptr[Nr] = d->nitems;
    cudaError_t e = cudaSetDevice(h->device);
    int32_t rc = e == cudaSuccess ? BBM_OK : bbm_fail(h, BBM_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    if (rc == BBM_OK) rc = upload(h, ents, &d->d_ents);
    if (rc == BBM_OK) rc = upload(h, items, &d->d_items);
    if (rc != BBM_OK) { bbm_sky_desc_destroy(d); return rc; }
    *out = d;
    return BBM_OK;
}

int32_t bbm_sky_desc_destroy(bbm_sky_desc_t d) {
    if (!d || d->magic != 0x5C71DE5Cu) return BBM_ERR_HANDLE;
    if (bbm_valid(d->h)) cudaSetDevice(d->h->device);
    if (d->d_ents) cudaFree(d->d_ents);
    if (d->d_items) cudaFree(d->d_items);
    for (int t = 0; t < 2; ++t) {
        if (d->tri[t].d_ents) cudaFree(d->tri[t].d_ents);
        if (d->tri[t].d_items) cudaFree(d->tri[t].d_items);
        if (d->tri[t].d_sub) cudaFree(d->tri[t].d_sub);
        if (d->tri[t].d_place) cudaFree(d->tri[t].d_place);
        if (d->tri[t].d_zero) cudaFree(d->tri[t].d_zero);
        for (auto& b : d->tri[t].inv_levels) free_batch(b);
        free_batch(d->tri[t].stepA);
        free_batch(d->tri[t].stepB);
        if (d->tri[t].flow_tiles) cudaFree(d->tri[t].flow_tiles);
        if (d->tri[t].flow_ctr) cudaFree(d->tri[t].flow_ctr);
        if (d->tri[t].d_cond) cudaFree(d->tri[t].d_cond);
    }
    if (d->rev) { bbm_sky_desc_destroy(d->rev); d->rev = nullptr; }
    if (d->d_titems) cudaFree(d->d_titems);
    if (d->d_struct) cudaFree(d->d_struct);
    if (d->qrf.d_steps) cudaFree(d->qrf.d_steps);
    free_batch(d->qrf.bW); free_batch(d->qrf.bC);
    if (d->qr.d_pan) cudaFree(d->qr.d_pan);
    if (d->qr.d_blk) cudaFree(d->qr.d_blk);
    if (d->qr.flow_tiles) cudaFree(d->qr.flow_tiles);
    if (d->qr.flow_ctr) cudaFree(d->qr.flow_ctr);
    free_batch(d->qr.gram); free_batch(d->qr.ytb); free_batch(d->qr.stepW); free_batch(d->qr.stepB);
    free_batch(d->rmul.batch);
    for (auto& b : d->qr.levels) free_batch(b);
    for (auto& kv : d->plans) free_plan(kv.second);
    d->magic = 0;
    delete d;
    return BBM_OK;
}

int32_t bbm_sky_desc_info(bbm_sky_desc_t d, int64_t* m, int64_t* n, int64_t* numentries, int64_t* nblocks) {
    if (!sky_valid(d)) return BBM_ERR_HANDLE;
    if (m) *m = d->m;
    if (n) *n = d->n;
    if (numentries) *numentries = d->numel;
    if (nblocks) *nblocks = d->nblocks;
    return BBM_OK;
}

int32_t bbm_sky_desc_layout(bbm_sky_desc_t d, int64_t* panel_offsets, int64_t* panel_strides, int64_t* klo, int64_t* khi) {
    if (!sky_valid(d)) return BBM_ERR_HANDLE;
    if (panel_offsets) std::copy(d->off.begin(), d->off.end(), panel_offsets);
    if (panel_strides) std::copy(d->S.begin(), d->S.end(), panel_strides);
    if (klo) std::copy(d->klo.begin(), d->klo.end(), klo);
    if (khi) std::copy(d->khi.begin(), d->khi.end(), khi);
    return BBM_OK;
}

int32_t bbm_sky_mul_f64(bbm_handle_t h, bbm_sky_desc_t d, double alpha, const double* d_data, const double* d_X, int64_t ldx,
                        int64_t nrhs, double beta, double* d_Y, int64_t ldy) {
    return sky_mul_impl<double>(h, d, alpha, d_data, d_X, ldx, nrhs, beta, d_Y, ldy);
}
int32_t bbm_sky_mul_f32(bbm_handle_t h, bbm_sky_desc_t d, float alpha, const float* d_data, const float* d_X, int64_t ldx,
                        int64_t nrhs, float beta, float* d_Y, int64_t ldy) {
    return sky_mul_impl<float>(h, d, alpha, d_data, d_X, ldx, nrhs, beta, d_Y, ldy);
}
int32_t bbm_sky_matmul_f64(bbm_handle_t h, bbm_sky_desc_t A, bbm_sky_desc_t B, bbm_sky_desc_t C, double alpha, const double* dA,
                           const double* dB, double beta, double* dC) {
    return sky_matmul_impl<double>(h, A, B, C, alpha, dA, dB, beta, dC);
}
int32_t bbm_sky_matmul_f32(bbm_handle_t h, bbm_sky_desc_t A, bbm_sky_desc_t B, bbm_sky_desc_t C, float alpha, const float* dA,
                           const float* dB, float beta, float* dC) {
    return sky_matmul_impl<float>(h, A, B, C, alpha, dA, dB, beta, dC);
}
int32_t bbm_sky_ql_f64(bbm_handle_t h, bbm_sky_desc_t d, double* d_data, double* d_tau) { return sky_ql_fact_impl(h, d, d_data, d_tau); }
int32_t bbm_sky_ql_lmul_adj_f64(bbm_handle_t h, bbm_sky_desc_t d, const double* d_factors, const double* d_tau, double* d_B, int64_t ldb,
                                int64_t nrhs) {
    return sky_ql_apply_impl(h, d, d_factors, d_tau, d_B, ldb, nrhs, 0);
}
int32_t bbm_sky_ql_ldiv_f64(bbm_handle_t h, bbm_sky_desc_t d, const double* d_factors, const double* d_tau, double* d_B, int64_t ldb,
                            int64_t nrhs) {
    return sky_ql_apply_impl(h, d, d_factors, d_tau, d_B, ldb, nrhs, 1);
}
int32_t bbm_sky_mul_t_f64(bbm_handle_t h, bbm_sky_desc_t d, double alpha, const double* d_data, const double* d_X, int64_t ldx,
                          int64_t nrhs, double beta, double* d_Y, int64_t ldy) {
    return sky_mul_t_impl<double>(h, d, alpha, d_data, d_X, ldx, nrhs, beta, d_Y, ldy);
}
int32_t bbm_sky_mul_t_f32(bbm_handle_t h, bbm_sky_desc_t d, float alpha, const float* d_data, const float* d_X, int64_t ldx,
                          int64_t nrhs, float beta, float* d_Y, int64_t ldy) {
    return sky_mul_t_impl<float>(h, d, alpha, d_data, d_X, ldx, nrhs, beta, d_Y, ldy);
}
int32_t bbm_sky_rmul_f64(bbm_handle_t h, bbm_sky_desc_t d, double alpha, const double* d_X, int64_t ldx, int64_t p, const double* d_data,
                         double beta, double* d_Y, int64_t ldy) {
    return sky_rmul_impl<double>(h, d, alpha, d_X, ldx, p, d_data, beta, d_Y, ldy);
}
int32_t bbm_sky_rmul_f32(bbm_handle_t h, bbm_sky_desc_t d, float alpha, const float* d_X, int64_t ldx, int64_t p, const float* d_data,
                         float beta, float* d_Y, int64_t ldy) {
    return sky_rmul_impl<float>(h, d, alpha, d_X, ldx, p, d_data, beta, d_Y, ldy);
}
int32_t bbm_sky_qr_f64(bbm_handle_t h, bbm_sky_desc_t d, double* d_data, double* d_tau) { return sky_qr_fact_impl(h, d, d_data, d_tau); }
int32_t bbm_sky_qr_lmul_adj_f64(bbm_handle_t h, bbm_sky_desc_t d, const double* d_factors, const double* d_tau, double* d_B, int64_t ldb,
                                int64_t nrhs) {
    return sky_qr_lmul_adj_impl(h, d, d_factors, d_tau, d_B, ldb, nrhs);
}
int32_t bbm_sky_qr_ldiv_f64(bbm_handle_t h, bbm_sky_desc_t d, const double* d_factors, const double* d_tau, double* d_B, int64_t ldb,
                            int64_t nrhs) {
    BBM_REQUIRE_HANDLE(h);
    if (!sky_valid(d) || d->h != h) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BlockSkylineMatrix descriptor");
    bool match = d->Nr == d->Nc;
    for (i64 K = 0; match && K < d->Nr; ++K) match = d->r[K] == d->c[K];
    if (!match) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "QR solve on the device needs a square matrix with matching diagonal blocks");
    int32_t rc = sky_qr_lmul_adj_impl(h, d, d_factors, d_tau, d_B, ldb, nrhs);
    if (rc != BBM_OK) return rc;
    return sky_trsm_impl<double>(h, d, 'U', 0, d_factors, d_B, ldb, nrhs);
}
int32_t bbm_sky_qr_f32(bbm_handle_t h, bbm_sky_desc_t d, float* d_data, float* d_tau) { return sky_fact_f32(h, d, d_data, d_tau, 0); }
int32_t bbm_sky_ql_f32(bbm_handle_t h, bbm_sky_desc_t d, float* d_data, float* d_tau) { return sky_fact_f32(h, d, d_data, d_tau, 1); }
int32_t bbm_sky_qr_lmul_adj_f32(bbm_handle_t h, bbm_sky_desc_t d, const float* d_factors, const float* d_tau, float* d_B, int64_t ldb, int64_t nrhs) {
    return sky_apply_f32(h, d, d_factors, d_tau, d_B, ldb, nrhs, 0);
}
int32_t bbm_sky_qr_ldiv_f32(bbm_handle_t h, bbm_sky_desc_t d, const float* d_factors, const float* d_tau, float* d_B, int64_t ldb, int64_t nrhs) {
    return sky_apply_f32(h, d, d_factors, d_tau, d_B, ldb, nrhs, 1);
}
int32_t bbm_sky_ql_lmul_adj_f32(bbm_handle_t h, bbm_sky_desc_t d, const float* d_factors, const float* d_tau, float* d_B, int64_t ldb, int64_t nrhs) {
    return sky_apply_f32(h, d, d_factors, d_tau, d_B, ldb, nrhs, 2);
}
int32_t bbm_sky_ql_ldiv_f32(bbm_handle_t h, bbm_sky_desc_t d, const float* d_factors, const float* d_tau, float* d_B, int64_t ldb, int64_t nrhs) {
    return sky_apply_f32(h, d, d_factors, d_tau, d_B, ldb, nrhs, 3);
}
int32_t bbm_sky_trsm_f64(bbm_handle_t h, bbm_sky_desc_t d, int32_t uplo, int32_t unit, const double* d_data, double* d_B,
                         int64_t ldb, int64_t nrhs) {
    return sky_trsm_impl<double>(h, d, uplo, unit, d_data, d_B, ldb, nrhs);
}
int32_t bbm_sky_trsm_f32(bbm_handle_t h, bbm_sky_desc_t d, int32_t uplo, int32_t unit, const float* d_data, float* d_B,
                         int64_t ldb, int64_t nrhs) {
    return sky_trsm_impl<float>(h, d, uplo, unit, d_data, d_B, ldb, nrhs);
}

}  // extern "C"
