/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the block-banded mul!/ldiv! hot path.
 *
 * A restatement, in plain C, of what BlockBandedMatrices.jl's CPU path computes for
 *   mul!(y, A::BandedBlockBandedMatrix, x)      
 *   mul!(y, A::BlockSkylineMatrix, x) / A*X     
 *   mul!(C, A::BlockSkylineMatrix, B::...)      
 *   ldiv!(Upper/LowerTriangular(A), B)          
 * and of the callers either side of it: A'x for both formats, A*B in band storage,
 * the triangular BandedBlockBandedMatrix substitution, qr!/ql! with lmul!(Q',B) and F \ B
 * (src/blockskylineqr.jl),
 * i.e. the [upstream BlockArrays/ArrayLayouts/BandedMatrices] block loops, whose bounds and
 * operands come from the reference files cited at each function, calling one BLAS routine per
 * block (gbmv / gemv / gemm / trsv).  The BLAS routines are either the small reference-order
 * loops below (default) or OpenBLAS's cblas entry points resolved at run time with
 * oracle_set_blas() (the same BLAS family Julia ships; used for the CPU baseline timing).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product never links or calls it.
 *
 * Parity status: layout pinned by the reference's integer goldens (tests/test_oracle_golden.py);
 * arithmetic pinned to "structured ~= dense", which is all the reference's own tests pin; the QR/QL
 * factors and taus are pinned to LAPACK's dense geqrf, as test/test_blockskylineqr.jl pins them.
 */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

enum { ORACLE_OK = 0, ORACLE_ERR_ARG = 1, ORACLE_ERR_DIM = 2, ORACLE_ERR_BAND = 3, ORACLE_ERR_BLAS = 7 };

/* ---- optional OpenBLAS binding (cblas, LP64) ------------------------------------------- */
enum { CblasColMajor = 102, CblasNoTrans = 111, CblasUpper = 121, CblasLower = 122,
       CblasNonUnit = 131, CblasUnit = 132 };

typedef void (*dgbmv_t)(int, int, int, int, int, int, double, const double*, int, const double*, int, double, double*, int);
typedef void (*sgbmv_t)(int, int, int, int, int, int, float, const float*, int, const float*, int, float, float*, int);
typedef void (*dgemv_t)(int, int, int, int, double, const double*, int, const double*, int, double, double*, int);
typedef void (*sgemv_t)(int, int, int, int, float, const float*, int, const float*, int, float, float*, int);
typedef void (*dgemm_t)(int, int, int, int, int, int, double, const double*, int, const double*, int, double, double*, int);
typedef void (*sgemm_t)(int, int, int, int, int, int, float, const float*, int, const float*, int, float, float*, int);
typedef void (*dtrsv_t)(int, int, int, int, int, const double*, int, double*, int);
typedef void (*strsv_t)(int, int, int, int, int, const float*, int, float*, int);
typedef void (*setthreads_t)(int);

static void* g_blas = NULL;
static dgbmv_t p_dgbmv; static sgbmv_t p_sgbmv;
static dgemv_t p_dgemv; static sgemv_t p_sgemv;
static dgemm_t p_dgemm; static sgemm_t p_sgemm;
static dtrsv_t p_dtrsv; static strsv_t p_strsv;
static setthreads_t p_setthreads;

static void* sym2(void* h, const char* a, const char* b) {
    void* p = dlsym(h, a);
    return p ? p : dlsym(h, b);
}

/* Bind to an OpenBLAS shared object (scipy's bundled build prefixes symbols with scipy_). */
int oracle_set_blas(const char* path) {
    if (!path) { g_blas = NULL; return ORACLE_OK; }
    void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return ORACLE_ERR_BLAS;
    p_dgbmv = (dgbmv_t)sym2(h, "scipy_cblas_dgbmv", "cblas_dgbmv");
    p_sgbmv = (sgbmv_t)sym2(h, "scipy_cblas_sgbmv", "cblas_sgbmv");
    p_dgemv = (dgemv_t)sym2(h, "scipy_cblas_dgemv", "cblas_dgemv");
    p_sgemv = (sgemv_t)sym2(h, "scipy_cblas_sgemv", "cblas_sgemv");
    p_dgemm = (dgemm_t)sym2(h, "scipy_cblas_dgemm", "cblas_dgemm");
    p_sgemm = (sgemm_t)sym2(h, "scipy_cblas_sgemm", "cblas_sgemm");
    p_dtrsv = (dtrsv_t)sym2(h, "scipy_cblas_dtrsv", "cblas_dtrsv");
    p_strsv = (strsv_t)sym2(h, "scipy_cblas_strsv", "cblas_strsv");
    p_setthreads = (setthreads_t)sym2(h, "scipy_openblas_set_num_threads", "openblas_set_num_threads");
    if (!p_dgbmv || !p_sgbmv || !p_dgemv || !p_sgemv || !p_dgemm || !p_sgemm || !p_dtrsv || !p_strsv)
        return ORACLE_ERR_BLAS;
    g_blas = h;
    return ORACLE_OK;
}

int oracle_blas_bound(void) { return g_blas != NULL; }

void oracle_set_blas_threads(int n) { if (g_blas && p_setthreads) p_setthreads(n); }

/* ---- index helpers shared by both precisions -------------------------------------------- */
static i64* prefix_sum(i64 n, const i64* s) {
    i64* p = (i64*)malloc(sizeof(i64) * (size_t)(n + 1));
    p[0] = 0;
    for (i64 i = 0; i < n; ++i) p[i + 1] = p[i] + s[i];
    return p;
}
static inline i64 imax(i64 a, i64 b) { return a > b ? a : b; }
static inline i64 imin(i64 a, i64 b) { return a < b ? a : b; }

/* Skyline panel table: src/BlockSkylineMatrix.jl:6-43,91-104. */
typedef struct { i64 Nr, Nc; i64 *R, *C, *off, *S, *klo, *khi; } sky_t;

static void sky_init(sky_t* s, i64 Nr, i64 Nc, const i64* r, const i64* c, const i64* l, const i64* u) {
    s->Nr = Nr; s->Nc = Nc;
    s->R = prefix_sum(Nr, r); s->C = prefix_sum(Nc, c);
    s->off = (i64*)malloc(sizeof(i64) * (size_t)(Nc + 1));
    s->S = (i64*)malloc(sizeof(i64) * (size_t)(Nc + 1));
    s->klo = (i64*)malloc(sizeof(i64) * (size_t)(Nc + 1));
    s->khi = (i64*)malloc(sizeof(i64) * (size_t)(Nc + 1));
    s->off[0] = 0;
    for (i64 J = 0; J < Nc; ++J) {
        i64 lo = imax(0, J - u[J]), hi = imin(Nr - 1, J + l[J]);
        s->klo[J] = lo; s->khi[J] = hi;
        s->S[J] = (hi >= lo) ? s->R[hi + 1] - s->R[lo] : 0;
        s->off[J + 1] = s->off[J] + s->S[J] * c[J];
    }
}
static void sky_free(sky_t* s) { free(s->R); free(s->C); free(s->off); free(s->S); free(s->klo); free(s->khi); }
static inline int sky_has(const sky_t* s, i64 K, i64 J) { return K >= s->klo[J] && K <= s->khi[J]; }
static inline i64 sky_start(const sky_t* s, i64 K, i64 J) { return s->off[J] + s->R[K] - s->R[s->klo[J]]; }

i64 oracle_sky_numentries(i64 Nr, i64 Nc, const i64* r, const i64* c, const i64* l, const i64* u) {
    sky_t s; sky_init(&s, Nr, Nc, r, c, l, u);
    i64 n = s.off[Nc];
    sky_free(&s);
    return n;
}

#define T double
#define SUF(name) name##_f64
#define P_GBMV p_dgbmv
#define P_GEMV p_dgemv
#define P_GEMM p_dgemm
#define P_TRSV p_dtrsv
#include "bbm_oracle_impl.inc"
#undef T
#undef SUF
#undef P_GBMV
#undef P_GEMV
#undef P_GEMM
#undef P_TRSV

#define T float
#define SUF(name) name##_f32
#define P_GBMV p_sgbmv
#define P_GEMV p_sgemv
#define P_GEMM p_sgemm
#define P_TRSV p_strsv
#include "bbm_oracle_impl.inc"
