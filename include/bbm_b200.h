/*
 * bbm_b200.h -- C ABI of libbbm_b200.so: hand-written sm_100a CUDA kernels for the data-parallel
 * hot path of BlockBandedMatrices.jl (mul!/muladd! for BandedBlockBandedMatrix and
 * BlockBandedMatrix/BlockSkylineMatrix, block-banded triangular ldiv!) and the callers either side of
 * it: triangular banded-block-banded products and solves, transposed products, A*B in band storage,
 * axpy!/lmul! on storage, sparse(A), qr!/ql! with lmul!(Q', B) and F \ B, and the row-block-partitioned
 * multi-GPU matvec.
 *
 * This is what a Julia package extension `ccall`s (INTEGRATION.md shows the binding): plain
 * pointers and sizes, no CUDA.jl / torch types.  All matrices keep the reference's storage
 * formats with shapes unchanged (citations are relative to the reference repository):
 *
 *   BandedBlockBandedMatrix  `data` = dense column-major P x n array, P = (l+u+1)(lam+mu+1),
 *       A[R[K]+k, C[J]+j] = data[(u+K-J)(lam+mu+1) + (mu+k-j) + P*(C[J]+j)]
 *       (src/BandedBlockBandedMatrix.jl:25-50, 503-506, 538-545; docs/src/index.md:53-84)
 *   BlockSkylineMatrix       `data` = vector of column-block-contiguous dense panels
 *       (src/BlockSkylineMatrix.jl:6-43, 93-104, 459-479; docs/src/index.md:28-50)
 *
 * Conventions
 *   - every entry point returns an int32 status (bbm_status); 0 = success.  The message of the
 *     last failure on a handle is bbm_last_error(h).  Nothing throws or aborts across the ABI.
 *     Julia maps BBM_ERR_DIM -> DimensionMismatch (where [upstream] BlockArrays' block loop and
 *     src/linalg.jl:42,61 throw), BBM_ERR_BAND -> BandError (src/BlockSkylineMatrix.jl:491).
 *   - block sizes / indices are int64 and 0-based; block sizes may be 0, bandwidths negative.
 *   - device pointers named d_* must be device memory of the handle's device; the caller owns
 *     every buffer; descriptors own only their index tables.
 *   - work is enqueued on the handle's stream and the call returns (asynchronous), except the
 *     *_host_* convenience calls and bbm_synchronize.
 *   - a handle is not thread-safe; distinct handles are independent.  Every call does its own
 *     cudaSetDevice, so calls may come from any host thread / Julia task.
 *   - the band-walk matvec moves `data` and `x` with 16-byte-granular bulk copies: the 16-byte-aligned
 *     windows around the addressed ranges must be readable, i.e. up to 15 bytes before the first and after the
 *     last addressed element (bytes outside the addressed range are never used).  bbm_malloc pads every
 *     allocation by 32 bytes for this; cudaMalloc memory satisfies it through its 256-byte granularity; a buffer
 *     carved out of a caller's own pool must leave 16 readable bytes on either side of `data` and `x`.
 *   - there is NO CPU fallback: without a usable CUDA device bbm_create fails with BBM_ERR_CUDA.
 */
#ifndef BBM_B200_H
#define BBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bbm_context* bbm_handle_t;
typedef struct bbm_bbb_desc* bbm_bbb_desc_t;
typedef struct bbm_sky_desc* bbm_sky_desc_t;

typedef enum bbm_status {
    BBM_OK = 0,
    BBM_ERR_HANDLE = 1,      /* null / destroyed handle or descriptor                          */
    BBM_ERR_DIM = 2,         /* DimensionMismatch: vector lengths, leading dims, block axes     */
    BBM_ERR_BAND = 3,        /* BandError: result structure too narrow, missing diagonal block  */
    BBM_ERR_ARG = 4,         /* invalid argument (null pointer, negative size, bad enum)        */
    BBM_ERR_CUDA = 5,        /* CUDA runtime error (message has the CUDA error string)          */
    BBM_ERR_NCCL = 6,        /* peer / collective setup error                                   */
    BBM_ERR_ALLOC = 7,       /* host or device allocation failed                                */
    BBM_ERR_UNSUPPORTED = 8  /* valid in the reference but not taken by the device path         */
} bbm_status;

/* ---- library / handle ------------------------------------------------------------------- */
int32_t bbm_version(void);                       /* major*10000 + minor*100 + patch             */
const char* bbm_status_string(int32_t status);
int32_t bbm_create(int32_t device, bbm_handle_t* out);
int32_t bbm_destroy(bbm_handle_t h);
int32_t bbm_set_stream(bbm_handle_t h, void* cuda_stream /* cudaStream_t, NULL = default */);
int32_t bbm_synchronize(bbm_handle_t h);
const char* bbm_last_error(bbm_handle_t h);
/* tuning / diagnostics knobs ("bbb.tile", "bbb.stages", "bbb.waves", "bbb.items_per_sm", "bbb.maxseg",
 * "bbb.l2hint", "bbb.kernel": 0 auto, 1 band-walk, 2 generic; "dist.graph": 1 replays a CUDA graph
 * of the whole NCCL-path multi-GPU step (default 0: measured slower); "dist.peer": 0 disables the
 * fused peer-memory halo push after bbm_dist_bbb_enable_peer); unknown keys -> BBM_ERR_ARG.       */
int32_t bbm_set_option(bbm_handle_t h, const char* key, int64_t value);
int32_t bbm_get_option(bbm_handle_t h, const char* key, int64_t* value);
/* number of kernels this handle has launched since creation (bench.py's gpu_launches)         */
int32_t bbm_launch_count(bbm_handle_t h, int64_t* count);

/* ---- memory utilities (so the Julia side needs no CUDA.jl) ------------------------------- */
int32_t bbm_malloc(bbm_handle_t h, void** d_ptr, size_t bytes);
int32_t bbm_free(bbm_handle_t h, void* d_ptr);
int32_t bbm_host_alloc(bbm_handle_t h, void** h_ptr, size_t bytes);   /* pinned host memory    */
int32_t bbm_host_free(bbm_handle_t h, void* h_ptr);
int32_t bbm_memcpy_h2d(bbm_handle_t h, void* d_dst, const void* h_src, size_t bytes);
int32_t bbm_memcpy_d2h(bbm_handle_t h, void* h_dst, const void* d_src, size_t bytes);
int32_t bbm_memcpy_d2d(bbm_handle_t h, void* d_dst, const void* d_src, size_t bytes);
int32_t bbm_memset(bbm_handle_t h, void* d_ptr, int32_t byte_value, size_t bytes);

/* ---- level-1 operations on storage arrays -------------------------------------------------
 * axpy!(a, A, B) for two matrices of IDENTICAL structure is y <- a*x + y on their `data`
 * (src/BandedBlockBandedMatrix.jl:605-606, src/broadcast.jl:307-314); lmul!(a, A) / rmul!(A, a) scale
 * `data` (src/BandedBlockBandedMatrix.jl:609-612).  n = number of stored entries (P*n resp. bb_numentries).
 * Operands of different bandwidths keep the reference's host path.                               */
int32_t bbm_axpy_f64(bbm_handle_t h, int64_t n, double alpha, const double* d_x, double* d_y);
int32_t bbm_axpy_f32(bbm_handle_t h, int64_t n, float alpha, const float* d_x, float* d_y);
int32_t bbm_scal_f64(bbm_handle_t h, int64_t n, double alpha, double* d_x);
int32_t bbm_scal_f32(bbm_handle_t h, int64_t n, float alpha, float* d_x);

/* ---- structured element-wise operations ------------------------------------------------------
 * C <- alpha*A + beta*B for operands with the SAME block axes but different bandwidths / storage formats: the
 * block-wise `A + B`, `A - B` (alpha, beta = 1, +-1; src/broadcast.jl:317-372: joint block range op(A,B), ranges only
 * one operand covers that operand, the rest of C's band zero), `C .= alpha .* A .+ B` (src/broadcast.jl:379-388 on
 * top of blockbanded_copyto! / blockbanded_axpy!, :307-314) and the copy of a matrix into a wider structure
 * (B = NULL; `BlockSkylineMatrix{T}(A, sizes)` of copy_accommodating_diagonals, src/BlockSkylineMatrix.jl:530).
 * The result structure (`similar`: the maximum of the bandwidths, src/broadcast.jl:391-411) stays on the host.
 * A or B may be absent (NULL descriptor).  BBM_ERR_DIM if block axes differ, BBM_ERR_BAND if an operand's bands do
 * not fit C.  d_C may alias an operand only if that operand has C's exact structure.  alpha*a and beta*b are rounded
 * separately and then added -- exactly a+b / a-b for unit coefficients, the un-fused `a .* x .+ y` otherwise.
 * bbm_bbb_add_*: all three in band storage.  bbm_sky_add_*: C block-skyline, each operand EITHER block-skyline
 * (X_sky) or banded-block-banded (X_bbb) -- the mixed sums of test/test_broadcasting.jl:205-215.               */
int32_t bbm_bbb_add_f64(bbm_handle_t h, bbm_bbb_desc_t A, bbm_bbb_desc_t B, bbm_bbb_desc_t C, double alpha, const double* d_A,
                        double beta, const double* d_B, double* d_C);
int32_t bbm_bbb_add_f32(bbm_handle_t h, bbm_bbb_desc_t A, bbm_bbb_desc_t B, bbm_bbb_desc_t C, float alpha, const float* d_A,
                        float beta, const float* d_B, float* d_C);
int32_t bbm_sky_add_f64(bbm_handle_t h, bbm_sky_desc_t A_sky, bbm_bbb_desc_t A_bbb, bbm_sky_desc_t B_sky, bbm_bbb_desc_t B_bbb,
                        bbm_sky_desc_t C, double alpha, const double* d_A, double beta, const double* d_B, double* d_C);
int32_t bbm_sky_add_f32(bbm_handle_t h, bbm_sky_desc_t A_sky, bbm_bbb_desc_t A_bbb, bbm_sky_desc_t B_sky, bbm_bbb_desc_t B_bbb,
                        bbm_sky_desc_t C, float alpha, const float* d_A, float beta, const float* d_B, float* d_C);
/* In place on the storage: A[i,i] += alpha*(lambda + d[i]), A[i,i+1] += alpha*du[i], A[i+1,i] += alpha*dl[i]; d_d, d_du,
 * d_dl may each be NULL (absent).  The second half of `A +- I`, `A +- Diagonal / Bidiagonal / Tridiagonal /
 * SymTridiagonal` (src/BlockSkylineMatrix.jl:535-638) and of `I - dt*Delta` for a BandedBlockBandedMatrix
 * (examples/finitedifference_2d.jl:50-51).  Needs square diagonal blocks (checksquareblocks, BBM_ERR_DIM) and
 * every touched entry stored (BBM_ERR_BAND otherwise, where the reference's setindex! throws).                */
int32_t bbm_bbb_add_diags_f64(bbm_handle_t h, bbm_bbb_desc_t d, double* d_data, double alpha, double lambda, const double* d_dl,
                              const double* d_d, const double* d_du);
int32_t bbm_bbb_add_diags_f32(bbm_handle_t h, bbm_bbb_desc_t d, float* d_data, float alpha, float lambda, const float* d_dl,
                              const float* d_d, const float* d_du);
int32_t bbm_sky_add_diags_f64(bbm_handle_t h, bbm_sky_desc_t d, double* d_data, double alpha, double lambda, const double* d_dl,
                              const double* d_d, const double* d_du);
int32_t bbm_sky_add_diags_f32(bbm_handle_t h, bbm_sky_desc_t d, float* d_data, float alpha, float lambda, const float* d_dl,
                              const float* d_d, const float* d_du);

/* ---- BandedBlockBandedMatrix -------------------------------------------------------------
 * Descriptor = the host-side fields of the reference struct (raxis, column axis, l, u, lam, mu;
 * src/BandedBlockBandedMatrix.jl:25-40) plus cached device index tables and the work
 * partition of the band-walk kernel.  r[Nr], c[Nc] are block sizes.                           */
int32_t bbm_bbb_desc_create(bbm_handle_t h, int64_t Nr, int64_t Nc, const int64_t* r, const int64_t* c,
                            int64_t l, int64_t u, int64_t lam, int64_t mu, bbm_bbb_desc_t* out);
int32_t bbm_bbb_desc_destroy(bbm_bbb_desc_t d);
/* m = sum r, n = sum c, P = rows of `data` (0 if -l>u or -lam>mu), work items of the partition */
int32_t bbm_bbb_desc_info(bbm_bbb_desc_t d, int64_t* m, int64_t* n, int64_t* P, int64_t* work_items);

/* Y <- alpha*A*X + beta*Y  (replaces [upstream BlockArrays] materialize!(::MatMulVecAdd/
 * ::MatMulMatAdd{<:AbstractBlockLayout,...}) + one BLAS gbmv per block, reached through
 * src/BandedBlockBandedMatrix.jl:263-269,405,514-545 and src/AbstractBlockBandedMatrix.jl:90-104;
 * in-repo rendering examples/cuarrays.jl:105-127).
 * d_data: P x n column-major, ld = P.  d_X: n x nrhs (ldx >= n), d_Y: m x nrhs (ldy >= m).
 * beta == 0 overwrites Y without reading it (test/test_blockskyline.jl:34-39).  Unused slots
 * of d_data are never read into the arithmetic (they may hold NaN).  Y must not alias X/data.  */
int32_t bbm_bbb_mul_f64(bbm_handle_t h, bbm_bbb_desc_t d, double alpha, const double* d_data,
                        const double* d_X, int64_t ldx, int64_t nrhs, double beta, double* d_Y, int64_t ldy);
int32_t bbm_bbb_mul_f32(bbm_handle_t h, bbm_bbb_desc_t d, float alpha, const float* d_data,
                        const float* d_X, int64_t ldx, int64_t nrhs, float beta, float* d_Y, int64_t ldy);

/* Y <- alpha*T*X + beta*Y with T = UpperTriangular(A) ('U') / LowerTriangular(A) ('L'), unit != 0 for the
 * Unit variants: the triangular BandedBlockBandedMatrix product of test/test_triblockbanded.jl:17-53
 * (layout TriangularLayout{UPLO,UNIT,BandedBlockBandedColumnMajor}; block bandwidths clipped as in
 * src/triblockbanded.jl:10-25).  Same kernels with a triangular mask: the other triangle of `data` is
 * never read into the arithmetic.  Needs matching square diagonal blocks, else BBM_ERR_UNSUPPORTED.  */
int32_t bbm_bbb_trimul_f64(bbm_handle_t h, bbm_bbb_desc_t d, int32_t uplo, int32_t unit, double alpha,
                           const double* d_data, const double* d_X, int64_t ldx, int64_t nrhs, double beta,
                           double* d_Y, int64_t ldy);
int32_t bbm_bbb_trimul_f32(bbm_handle_t h, bbm_bbb_desc_t d, int32_t uplo, int32_t unit, float alpha,
                           const float* d_data, const float* d_X, int64_t ldx, int64_t nrhs, float beta,
                           float* d_Y, int64_t ldy);

/* Y <- alpha*A^T*X + beta*Y: `mul!(y, A', x)` / `transpose(A)*x` for real element types (the adjoint's
 * Rows layouts and reversed bandwidths: src/adjtransblockbanded.jl:2-7, src/AbstractBlockBandedMatrix.jl:32-40;
 * [upstream] block loop + gbmv 'T' per block on the same band arrays).  d_X: m x nrhs, d_Y: n x nrhs.      */
int32_t bbm_bbb_mul_t_f64(bbm_handle_t h, bbm_bbb_desc_t d, double alpha, const double* d_data, const double* d_X,
                          int64_t ldx, int64_t nrhs, double beta, double* d_Y, int64_t ldy);
int32_t bbm_bbb_mul_t_f32(bbm_handle_t h, bbm_bbb_desc_t d, float alpha, const float* d_data, const float* d_X,
                          int64_t ldx, int64_t nrhs, float beta, float* d_Y, int64_t ldy);

/* C <- alpha*A*B + beta*C, all three BandedBlockBandedMatrix (replaces [upstream] _block_muladd! with one
 * banded x banded gbmm per block triple; `similar(::MulAdd)` of src/linalg.jl:50-71 -- bandwidths
 * (lA+lB, uA+uB), (lamA+lamB, muA+muB) -- stays on the host; test/test_linalg.jl:91-104, 160-171).
 * BBM_ERR_DIM for incompatible block axes, BBM_ERR_BAND if the product's bands do not fit C.      */
int32_t bbm_bbb_matmul_f64(bbm_handle_t h, bbm_bbb_desc_t A, bbm_bbb_desc_t B, bbm_bbb_desc_t C, double alpha,
                           const double* d_A, const double* d_B, double beta, double* d_C);
int32_t bbm_bbb_matmul_f32(bbm_handle_t h, bbm_bbb_desc_t A, bbm_bbb_desc_t B, bbm_bbb_desc_t C, float alpha,
                           const float* d_A, const float* d_B, float beta, float* d_C);

/* B <- T^-1 B in place, T = Upper ('U') / LowerTriangular ('L') of a BandedBlockBandedMatrix, unit != 0 for the
 * Unit variants: `ldiv!(LowerTriangular(A), b)`, `Ldiv(U, b)` of test/test_triblockbanded.jl:80-99 and the
 * Gauss-Seidel sweep of examples/finitedifference_2d.jl:17-47 (replaces [upstream] block substitution with a
 * banded tbsv per diagonal block and a gbmv per off-diagonal block).  One thread-block cluster runs a systolic
 * level schedule with the recent solution values in (distributed) shared memory.  Needs matching square
 * diagonal blocks and non-negative bandwidths; BBM_ERR_UNSUPPORTED beyond about 24k block rows x small bands.
 * d_B: m x nrhs, ldb >= m; right-hand sides are solved concurrently, one thread-block cluster each (batches of 16).  */
int32_t bbm_bbb_trsv_f64(bbm_handle_t h, bbm_bbb_desc_t d, int32_t uplo, int32_t unit, const double* d_data,
                         double* d_B, int64_t ldb, int64_t nrhs);
int32_t bbm_bbb_trsv_f32(bbm_handle_t h, bbm_bbb_desc_t d, int32_t uplo, int32_t unit, const float* d_data,
                         float* d_B, int64_t ldb, int64_t nrhs);

/* The same solve split in two for repeated use with ONE matrix -- the Gauss-Seidel loop of examples/finitedifference_2d.jl:17-47
 * solves with the same LowerTriangular(A) in every sweep: prepare gathers the matrix entries into the wave-ordered planes the
 * solve kernel streams (a plan-owned buffer of about (1 + (l or u)(lam+mu+1) + (lam or mu)) x 1.5 x m entries), solve only moves
 * the right-hand sides.  Several right-hand sides run concurrently, one thread-block cluster each.  The plan stays valid while
 * d_data is unchanged; BBM_ERR_UNSUPPORTED if the wave-ordered kernels do not apply (use bbm_bbb_trsv_* then).             */
typedef struct bbm_bbb_tri_plan* bbm_bbb_tri_plan_t;
int32_t bbm_bbb_trsv_prepare_f64(bbm_handle_t h, bbm_bbb_desc_t d, int32_t uplo, int32_t unit, const double* d_data, bbm_bbb_tri_plan_t* out);
int32_t bbm_bbb_trsv_prepare_f32(bbm_handle_t h, bbm_bbb_desc_t d, int32_t uplo, int32_t unit, const float* d_data, bbm_bbb_tri_plan_t* out);
int32_t bbm_bbb_trsv_solve_f64(bbm_bbb_tri_plan_t p, double* d_B, int64_t ldb, int64_t nrhs);
int32_t bbm_bbb_trsv_solve_f32(bbm_bbb_tri_plan_t p, float* d_B, int64_t ldb, int64_t nrhs);
int32_t bbm_bbb_tri_plan_destroy(bbm_bbb_tri_plan_t p);

/* sparse(A) (ext/BlockBandedMatricesSparseArraysExt.jl:10-27) as CSC arrays built on the device: every in-band entry
 * of every in-band block, explicit zeros included, rows ascending within a column.  Two calls, like any CSC build:
 *   bbm_bbb_sparse_colptr  fills d_colptr (n+1 entries) and returns nnz (synchronises for that number; pass
 *                          nnz = NULL to stay asynchronous);
 *   bbm_bbb_sparse_fill_*  writes d_rowval / d_nzval (nnz entries each).
 * index_base = 1 gives Julia's SparseMatrixCSC(m, n, colptr, rowval, nzval) directly, 0 the C / cuSPARSE form.  */
int32_t bbm_bbb_sparse_colptr(bbm_handle_t h, bbm_bbb_desc_t d, int32_t index_base, int64_t* d_colptr, int64_t* nnz);
int32_t bbm_bbb_sparse_fill_f64(bbm_handle_t h, bbm_bbb_desc_t d, const double* d_data, const int64_t* d_colptr,
                                int32_t index_base, int64_t* d_rowval, double* d_nzval);
int32_t bbm_bbb_sparse_fill_f32(bbm_handle_t h, bbm_bbb_desc_t d, const float* d_data, const int64_t* d_colptr,
                                int32_t index_base, int64_t* d_rowval, float* d_nzval);

/* Same product with HOST vectors (the reference-facing call when x, y live in Julia Arrays and
 * the matrix is device-resident): copies X host->device, runs the kernel, copies Y back and
 * synchronises.  h_X / h_Y should be pinned (bbm_host_alloc) for full PCIe speed.             */
int32_t bbm_bbb_mul_host_f64(bbm_handle_t h, bbm_bbb_desc_t d, double alpha, const double* d_data,
                             const double* h_X, int64_t ldx, int64_t nrhs, double beta, double* h_Y, int64_t ldy);
int32_t bbm_bbb_mul_host_f32(bbm_handle_t h, bbm_bbb_desc_t d, float alpha, const float* d_data,
                             const float* h_X, int64_t ldx, int64_t nrhs, float beta, float* h_Y, int64_t ldy);

/* ---- BlockSkylineMatrix / BlockBandedMatrix ------------------------------------------------
 * Descriptor = BlockSkylineSizes (src/BlockSkylineMatrix.jl:47-53): block axes and per-column
 * block bandwidths lvec[Nc], uvec[Nc] (constant vectors for BlockBandedMatrix); it computes the
 * panel offsets / strides exactly as bb_blockstarts / bb_blockstrides / bb_numentries
 * (src/BlockSkylineMatrix.jl:6-43, 93-104) and caches the device work lists.                    */
typedef struct bbm_sky_desc* bbm_sky_desc_t;
int32_t bbm_sky_desc_create(bbm_handle_t h, int64_t Nr, int64_t Nc, const int64_t* r, const int64_t* c,
                            const int64_t* lvec, const int64_t* uvec, bbm_sky_desc_t* out);
int32_t bbm_sky_desc_destroy(bbm_sky_desc_t d);
int32_t bbm_sky_desc_info(bbm_sky_desc_t d, int64_t* m, int64_t* n, int64_t* numentries, int64_t* nblocks);
/* panel_offsets[Nc+1] (0-based offset of panel J in `data`), panel_strides[Nc] (= block_strides),
 * klo[Nc], khi[Nc] (0-based inclusive block-row range of column J; khi < klo if empty); any may
 * be NULL.  For cross-checking against the host container's block_starts / block_strides.        */
int32_t bbm_sky_desc_layout(bbm_sky_desc_t d, int64_t* panel_offsets, int64_t* panel_strides, int64_t* klo, int64_t* khi);

/* Y <- alpha*A*X + beta*Y, A block-skyline (replaces [upstream] materialize!(::MatMulVecAdd /
 * ::MatMulMatAdd{<:AbstractBlockLayout,...}) + one BLAS gemv/gemm per block on the operands of
 * src/BlockSkylineMatrix.jl:442-465; pinned by test/test_linalg.jl:49-57,
 * test/test_blockskyline.jl:28-39).  d_data: the skyline `data` vector.                          */
int32_t bbm_sky_mul_f64(bbm_handle_t h, bbm_sky_desc_t d, double alpha, const double* d_data, const double* d_X,
                        int64_t ldx, int64_t nrhs, double beta, double* d_Y, int64_t ldy);
int32_t bbm_sky_mul_f32(bbm_handle_t h, bbm_sky_desc_t d, float alpha, const float* d_data, const float* d_X,
                        int64_t ldx, int64_t nrhs, float beta, float* d_Y, int64_t ldy);

/* Y <- alpha * A' * X + beta * Y: mul!(y, A', x) / transpose(A)*x for a real BlockSkylineMatrix
 * (src/adjtransblockbanded.jl:2-7: the adjoint keeps a block-banded layout with reversed bandwidths).
 * d_X: m x nrhs (ldx >= m), d_Y: n x nrhs (ldy >= n); beta == 0 never reads Y.                         */
int32_t bbm_sky_mul_t_f64(bbm_handle_t h, bbm_sky_desc_t d, double alpha, const double* d_data, const double* d_X,
                          int64_t ldx, int64_t nrhs, double beta, double* d_Y, int64_t ldy);
int32_t bbm_sky_mul_t_f32(bbm_handle_t h, bbm_sky_desc_t d, float alpha, const float* d_data, const float* d_X,
                          int64_t ldx, int64_t nrhs, float beta, float* d_Y, int64_t ldy);

/* Y <- alpha * X * A + beta * Y: `X*A` / `similar(X) .= MulAdd(X, A)` for a dense column-major X (p x m, ldx >= p) and a
 * BlockSkylineMatrix A (test/test_linalg.jl:56-57; [upstream] block loop Y[:,J] += X[:,K] * A_KJ with one gemm per block).
 * d_Y: p x n (ldy >= p); beta == 0 never reads Y.  One dense product per block COLUMN of A (its in-band blocks are one
 * contiguous panel), batched into one launch.                                                              */
int32_t bbm_sky_rmul_f64(bbm_handle_t h, bbm_sky_desc_t d, double alpha, const double* d_X, int64_t ldx, int64_t p,
                         const double* d_data, double beta, double* d_Y, int64_t ldy);
int32_t bbm_sky_rmul_f32(bbm_handle_t h, bbm_sky_desc_t d, float alpha, const float* d_X, int64_t ldx, int64_t p,
                         const float* d_data, float beta, float* d_Y, int64_t ldy);

/* C <- alpha*A*B + beta*C, all three block-skyline (replaces [upstream] _block_muladd! + one BLAS
 * gemm per block triple; C's structure is the caller's: src/linalg.jl:27-48 for A*B, a wider C is
 * allowed, test/test_blockbanded.jl:213-221).  BBM_ERR_DIM if the block axes are incompatible,
 * BBM_ERR_BAND if a product block is not stored in C.  Float64 runs on the FP64 tensor cores
 * (DMMA); "sky.mm_kernel" = 2 forces the SIMT kernel.  C must not alias A or B.                  */
int32_t bbm_sky_matmul_f64(bbm_handle_t h, bbm_sky_desc_t A, bbm_sky_desc_t B, bbm_sky_desc_t C, double alpha,
                           const double* d_A, const double* d_B, double beta, double* d_C);
int32_t bbm_sky_matmul_f32(bbm_handle_t h, bbm_sky_desc_t A, bbm_sky_desc_t B, bbm_sky_desc_t C, float alpha,
                           const float* d_A, const float* d_B, float beta, float* d_C);

/* B <- T^-1 B in place, T = UpperTriangular ('U') / LowerTriangular ('L') of A, unit != 0 for the
 * Unit variants (replaces [upstream] materialize!(::Ldiv{<:TriangularLayout{..,<:AbstractBlockLayout}})
 * block substitution: BLAS trsv on the diagonal block + gemv on the off-diagonal panel of
 * src/linalg.jl:93-115, bounds src/triblockbanded.jl:10-25; callers src/blockskylineqr.jl:136-155).
 * Needs matching square diagonal blocks (hasmatchingblocks), else BBM_ERR_UNSUPPORTED -- the
 * reference takes a generic fallback there; BBM_ERR_BAND if a diagonal block is not stored.
 * d_B: m x nrhs, ldb >= m.  All right-hand sides are solved together.                            */
int32_t bbm_sky_trsm_f64(bbm_handle_t h, bbm_sky_desc_t d, int32_t uplo, int32_t unit, const double* d_data,
                         double* d_B, int64_t ldb, int64_t nrhs);
int32_t bbm_sky_trsm_f32(bbm_handle_t h, bbm_sky_desc_t d, int32_t uplo, int32_t unit, const float* d_data,
                         float* d_B, int64_t ldb, int64_t nrhs);

/* lmul!(F.Q', B) and ldiv!(F, B) for F = qr(A) of a BlockBandedMatrix (src/blockskylineqr.jl:77-127 and :131-145;
 * test/test_blockskylineqr.jl:29-34).  `d` describes F.factors -- block-skyline with bandwidths (l, l+u), R in and
 * above the diagonal blocks, the Householder vectors of panel K below them in block rows K..K+l (LAPACK packing) --
 * d_tau is F.tau (length n).  The factorisation itself stays where the reference computes it (qr! on the host, one
 * LAPACK panel after the other); these calls replace the per-panel ormqr + the triangular solve: every panel is
 * brought into compact-WY form in parallel, then two tensor-core products per panel run in sequence.
 * Float64; at least as many block rows as block columns; bbm_sky_qr_ldiv needs a square matrix with matching
 * diagonal blocks (BBM_ERR_UNSUPPORTED otherwise: the call falls through to the reference).  In place on B.  */
int32_t bbm_sky_qr_lmul_adj_f64(bbm_handle_t h, bbm_sky_desc_t d, const double* d_factors, const double* d_tau,
                                double* d_B, int64_t ldb, int64_t nrhs);
/* qr!(A) in place (_blockbanded_qr!, src/blockskylineqr.jl:29-49): d_data holds A in the skyline storage of the
 * FACTORS structure -- block bandwidths (l, l+u), the extra upper blocks zero, exactly what the reference's `_qr`
 * (:75) allocates -- and becomes F.factors; d_tau (length n for at least as many block rows as columns, else m)
 * becomes F.tau.  Same Householder convention as LinearAlgebra.reflector! / LAPACK geqrf, so factors and tau
 * agree with the reference's to rounding.  Uniform block bandwidths, Float64.                              */
int32_t bbm_sky_qr_f64(bbm_handle_t h, bbm_sky_desc_t d, double* d_data, double* d_tau);

/* The QL counterparts (ql!, lmul!(F.Q', B), ldiv!(F::QL, B); src/blockskylineqr.jl:51-72, 96-111, 149-157): `d`
 * describes the factors structure (l+u, u); factors and tau come out in LAPACK geqlf's packing (L in and below the
 * diagonal blocks, the Householder vectors above).  Implemented by reversal -- the row-and-column reversed matrix
 * is the flat skyline vector read backwards -- on top of the QR kernels.  Square block structures, Float64.   */
int32_t bbm_sky_ql_f64(bbm_handle_t h, bbm_sky_desc_t d, double* d_data, double* d_tau);
int32_t bbm_sky_ql_lmul_adj_f64(bbm_handle_t h, bbm_sky_desc_t d, const double* d_factors, const double* d_tau,
                                double* d_B, int64_t ldb, int64_t nrhs);
int32_t bbm_sky_ql_ldiv_f64(bbm_handle_t h, bbm_sky_desc_t d, const double* d_factors, const double* d_tau,
                            double* d_B, int64_t ldb, int64_t nrhs);
int32_t bbm_sky_qr_ldiv_f64(bbm_handle_t h, bbm_sky_desc_t d, const double* d_factors, const double* d_tau,
                            double* d_B, int64_t ldb, int64_t nrhs);
/* The Float32 counterparts of the six calls above (src/blockskylineqr.jl:17-67 is generic in T).  They run the Float64
 * kernels on widened copies (stream-ordered scratch: 2 x the factors' storage) and round once at the end, so factors, tau
 * and solutions agree with the reference's Float32 arithmetic to Float32 rounding; same status codes.                  */
int32_t bbm_sky_qr_f32(bbm_handle_t h, bbm_sky_desc_t d, float* d_data, float* d_tau);
int32_t bbm_sky_ql_f32(bbm_handle_t h, bbm_sky_desc_t d, float* d_data, float* d_tau);
int32_t bbm_sky_qr_lmul_adj_f32(bbm_handle_t h, bbm_sky_desc_t d, const float* d_factors, const float* d_tau, float* d_B, int64_t ldb, int64_t nrhs);
int32_t bbm_sky_qr_ldiv_f32(bbm_handle_t h, bbm_sky_desc_t d, const float* d_factors, const float* d_tau, float* d_B, int64_t ldb, int64_t nrhs);
int32_t bbm_sky_ql_lmul_adj_f32(bbm_handle_t h, bbm_sky_desc_t d, const float* d_factors, const float* d_tau, float* d_B, int64_t ldb, int64_t nrhs);
int32_t bbm_sky_ql_ldiv_f32(bbm_handle_t h, bbm_sky_desc_t d, const float* d_factors, const float* d_tau, float* d_B, int64_t ldb, int64_t nrhs);

/* ---- several GPUs: row-block partition + halo exchange ------------------------------------
 * One process per GPU.  Block rows are split into contiguous ranges (the reference's one
 * data-parallel strategy, examples/shared_array_backend.jl:92-122, partition :106-107); rank p owns
 * y[K0:K1) and x[J0:J1) and holds the column slab [Jlo,Jhi) of `data` that its rows touch -- a
 * block-row range of a BandedBlockBandedMatrix is banded-block-banded with bandwidths
 * (l-s, u+s) over a column range of the same storage (src/BandedBlockBandedMatrix.jl:416-420,
 * 450-456).  Only the x block-columns inside the l+u block halo move between GPUs.           */
typedef struct bbm_comm* bbm_comm_t;
typedef struct bbm_dist_bbb_plan* bbm_dist_bbb_plan_t;

typedef struct bbm_dist_layout {
    int64_t K0, K1;        /* owned block rows [K0,K1)                                           */
    int64_t J0, J1;        /* owned block columns (the part of x this rank owns)                 */
    int64_t Jlo, Jhi;      /* local block columns: hull of owned and touched columns             */
    int64_t Ka, Kb;        /* interior block rows [Ka,Kb): touch owned columns only              */
    int64_t m_own, n_own;  /* rows of y / entries of x owned                                      */
    int64_t n_local;       /* length of x_local = columns of the local data slab                 */
    int64_t col_offset;    /* first global column of the local slab (= C[Jlo])                   */
    int64_t own_offset;    /* offset of the owned x inside x_local                               */
    int64_t row_offset;    /* first global row owned (= R[K0])                                   */
    int64_t nsend, nrecv;  /* halo messages per exchange                                          */
} bbm_dist_layout;

/* Host-only planning (no GPU, no handle): balanced contiguous block-row ranges, Kbounds[nranks+1]. */
int32_t bbm_bbb_partition_rows(int64_t Nr, const int64_t* r, int32_t nranks, int64_t* Kbounds);
/* Host-only: layout of `rank`; sendlist / recvlist (may be NULL) receive nsend / nrecv triples
 * (peer, offset into x_local, element count); each needs room for 3*(nranks-1) entries.          */
int32_t bbm_bbb_dist_layout(int64_t Nr, int64_t Nc, const int64_t* r, const int64_t* c, int64_t l, int64_t u,
                            int32_t nranks, int32_t rank, const int64_t* Kbounds, bbm_dist_layout* out,
                            int64_t* sendlist, int64_t* recvlist);

/* NCCL communicator over the ranks (libnccl.so.2 is bound at run time).  Rank 0 calls
 * bbm_comm_unique_id and ships the 128 bytes to the others by any means (MPI.jl, Distributed,
 * torch.distributed, a file), then every rank calls bbm_comm_create (collective).
 * id128 == NULL with nranks > 1 makes a communicator without NCCL: the caller fills the halo
 * parts of x_local itself (its own MPI exchange) and bbm_dist_bbb_mul_* only multiplies.        */
int32_t bbm_comm_unique_id(uint8_t* id128);
int32_t bbm_comm_create(bbm_handle_t h, const uint8_t* id128, int32_t nranks, int32_t rank, bbm_comm_t* out);
int32_t bbm_comm_destroy(bbm_comm_t c);

/* Plan for one global matrix structure; Kbounds == NULL takes bbm_bbb_partition_rows.           */
int32_t bbm_dist_bbb_plan_create(bbm_comm_t c, int64_t Nr, int64_t Nc, const int64_t* r, const int64_t* cols,
                                 int64_t l, int64_t u, int64_t lam, int64_t mu, const int64_t* Kbounds,
                                 bbm_dist_bbb_plan_t* out);
int32_t bbm_dist_bbb_plan_destroy(bbm_dist_bbb_plan_t p);
int32_t bbm_dist_bbb_plan_layout(bbm_dist_bbb_plan_t p, bbm_dist_layout* out);

/* y_own <- alpha * A[K0:K1, :] * x + beta * y_own.
 * d_data_local: P x n_local column slab (global columns col_offset ... ) of `data`.
 * d_x_local:    n_local x nrhs (ld = max(1,n_local)); the caller fills the owned part at
 *               own_offset, the halo parts are overwritten by the exchange.
 * d_y_own:      m_own x nrhs (ld = max(1,m_own)).
 * The halo exchange (grouped ncclSend/ncclRecv on the communicator's stream) overlaps the
 * product of the interior block rows; boundary rows follow.  Collective over the ranks.        */
/* Optional, collective, once per plan: switch the halo exchange from NCCL send/recv to a FUSED
 * kernel.  Every rank gets a small device block (epoch flags + double-buffered halo buffers) that is
 * mapped into both neighbours (CUDA IPC over NVLink).  A pusher CTA of every rank's band-walk kernel
 * stores the rank's boundary x blocks straight into the neighbours' halo buffers and raises an epoch
 * flag; CTAs whose steps read a halo block column wait for it and take that x block from the buffer.
 * One launch per rank per step, no NCCL kernel, no host synchronisation.  Halos must come from the
 * immediate neighbours only, else BBM_ERR_UNSUPPORTED on every rank and the NCCL path stays in use.
 * On this path the halo parts of x_local are neither read nor written; nrhs must be 1.            */
int32_t bbm_dist_bbb_enable_peer(bbm_dist_bbb_plan_t p, int32_t elem_size);
/* active: fused path in use; timed_out: a peer wait expired in some call (results invalid).  Synchronises. */
int32_t bbm_dist_bbb_peer_status(bbm_dist_bbb_plan_t p, int32_t* active, int32_t* timed_out);

int32_t bbm_dist_bbb_mul_f64(bbm_dist_bbb_plan_t p, double alpha, const double* d_data_local, double* d_x_local,
                             int64_t nrhs, double beta, double* d_y_own);
int32_t bbm_dist_bbb_mul_f32(bbm_dist_bbb_plan_t p, float alpha, const float* d_data_local, float* d_x_local,
                             int64_t nrhs, float beta, float* d_y_own);

/* The same product with the rank's vectors on the HOST (the reference-facing call when x and y live in Julia Arrays):
 * h_x_own (n_own entries) is uploaded, the halo blocks travel over NCCL, the owned rows are multiplied and h_y_own
 * (m_own entries) comes back -- pipelined over block-row chunks (upload of chunk i+1, product of chunk i, download of
 * chunk i-1 and the halo exchange all overlap), synchronous like bbm_bbb_mul_host_*.  One right-hand side.  The device
 * workspaces belong to the plan.  Pinned host memory (bbm_host_alloc) for full PCIe speed.  Collective.            */
int32_t bbm_dist_bbb_mul_host_f64(bbm_dist_bbb_plan_t p, double alpha, const double* d_data_local, const double* h_x_own, double beta,
                                  double* h_y_own);
int32_t bbm_dist_bbb_mul_host_f32(bbm_dist_bbb_plan_t p, float alpha, const float* d_data_local, const float* h_x_own, float beta,
                                  float* h_y_own);

/* ---- several GPUs: row-block-partitioned BlockBandedMatrix (matvec and A*B, cfg4) -----------------------------
 * Block rows are split exactly like the banded-block-banded case (bbm_bbb_partition_rows; x / the block rows of a
 * right factor B are owned like y).  A block-row range [S0,S1) of a BlockBandedMatrix(rows, cols, (l,u)) is a
 * BlockSkylineMatrix over the block columns [Jlo,Jhi) it touches, with per-column bandwidths (cf. the sub-view
 * bandwidths of src/linalg.jl:78-87); its storage is NOT a slice of the parent's, so each rank holds its own panels
 * in exactly the layout bbm_sky_desc_create gives for (rows[S0:S1], cols[Jlo:Jhi], lvec, uvec).                      */
typedef struct bbm_dist_sky_plan* bbm_dist_sky_plan_t;
typedef struct bbm_dist_sky_mm_plan* bbm_dist_sky_mm_plan_t;

/* Host-only planning (no GPU, no handle).  lvec / uvec receive Jhi-Jlo entries (NULL to query the range only).      */
int32_t bbm_sky_rowrange_structure(int64_t Nr, int64_t Nc, int64_t l, int64_t u, int64_t S0, int64_t S1, int64_t* Jlo, int64_t* Jhi,
                                   int64_t* lvec, int64_t* uvec);
/* Host-only: ranges6 = {K0,K1 (owned block rows of A and C), S0,S1 (block rows of B stored locally), O0,O1 (block rows
 * of B owned)}; send_blocks[q] / recv_blocks[q] = number of halo blocks of B exchanged with rank q (length nranks).  */
int32_t bbm_dist_sky_mm_layout(int64_t NrA, int64_t NcA, int64_t NcB, const int64_t* cA, const int64_t* cB, int64_t lA, int64_t uA,
                               int64_t lB, int64_t uB, int32_t nranks, int32_t rank, const int64_t* Kbounds, int64_t* ranges6,
                               int64_t* send_blocks, int64_t* recv_blocks);

/* y_own <- alpha * A[K0:K1,:] * x + beta * y_own.  d_data_local: the rank's panels (numentries_local entries);
 * d_x_local: n_local x nrhs like bbm_dist_bbb_mul_* (owned part at own_offset, halo parts overwritten by the NCCL
 * exchange, which overlaps the product of the interior block rows); the local structure's columns start at x_offset
 * inside x_local.  Collective over the ranks.                                                                    */
int32_t bbm_dist_sky_plan_create(bbm_comm_t c, int64_t Nr, int64_t Nc, const int64_t* r, const int64_t* cols, int64_t l, int64_t u,
                                 const int64_t* Kbounds, bbm_dist_sky_plan_t* out);
int32_t bbm_dist_sky_plan_destroy(bbm_dist_sky_plan_t p);
int32_t bbm_dist_sky_plan_layout(bbm_dist_sky_plan_t p, bbm_dist_layout* out, int64_t* Jlo, int64_t* Jhi, int64_t* numentries_local,
                                 int64_t* x_offset);
int32_t bbm_dist_sky_plan_bandwidths(bbm_dist_sky_plan_t p, int64_t* lvec, int64_t* uvec);
int32_t bbm_dist_sky_mul_f64(bbm_dist_sky_plan_t p, double alpha, const double* d_data_local, double* d_x_local, int64_t nrhs, double beta,
                             double* d_y_own);
int32_t bbm_dist_sky_mul_f32(bbm_dist_sky_plan_t p, float alpha, const float* d_data_local, float* d_x_local, int64_t nrhs, float beta,
                             float* d_y_own);

/* C[K0:K1,:] <- alpha * A[K0:K1,:] * B + beta * C[K0:K1,:]  (cfg4: "row blocks sharded over 8 GPUs").  A_loc / C_loc hold the
 * rank's own block rows of A and of C (block bandwidths (lA+lB, uA+uB), `similar(::MulAdd)` of src/linalg.jl:27-48), B_loc
 * the block rows [S0,S1) of B the rank multiplies with; rows of B_loc outside the owned range [O0,O1) are halo: every call
 * packs the blocks the neighbours need (gather kernel), exchanges them with ncclSend/ncclRecv on the communicator's stream
 * and scatters them into d_B_local, while the tiles of C that only multiply owned rows already run on the FP64 tensor
 * cores; the remaining tiles follow the arrival.  which = 0 / 1 / 2 selects A_loc / B_loc / C_loc in the queries.   */
int32_t bbm_dist_sky_mm_plan_create(bbm_comm_t c, int64_t NrA, int64_t NcA, int64_t NcB, const int64_t* rA, const int64_t* cA,
                                    const int64_t* cB, int64_t lA, int64_t uA, int64_t lB, int64_t uB, const int64_t* Kbounds,
                                    bbm_dist_sky_mm_plan_t* out);
int32_t bbm_dist_sky_mm_plan_destroy(bbm_dist_sky_mm_plan_t p);
int32_t bbm_dist_sky_mm_plan_info(bbm_dist_sky_mm_plan_t p, int64_t* ranges6, int64_t* Jlo3, int64_t* Jhi3, int64_t* numentries3);
int32_t bbm_dist_sky_mm_plan_bandwidths(bbm_dist_sky_mm_plan_t p, int32_t which, int64_t* lvec, int64_t* uvec);
int32_t bbm_dist_sky_matmul_f64(bbm_dist_sky_mm_plan_t p, double alpha, const double* d_A_local, double* d_B_local, double beta,
                                double* d_C_local);
int32_t bbm_dist_sky_matmul_f32(bbm_dist_sky_mm_plan_t p, float alpha, const float* d_A_local, float* d_B_local, float beta,
                                float* d_C_local);

#ifdef __cplusplus
}
#endif
#endif /* BBM_B200_H */
