// sparse(A) for a BandedBlockBandedMatrix (ext/BlockBandedMatricesSparseArraysExt.jl:10-27): the reference
// collects (row, column, value) triplets block by block on the host and lets SparseArrays sort them.  In band
// storage a column of `data` already holds the column's in-band entries in ascending row order (slab = block row,
// band row = in-block row), so the CSC form is a per-column compaction of the valid slots:
//   pass 1  nnz per column (pure index arithmetic) -> exclusive scan -> colptr
//   pass 2  one thread per column copies its valid slots and their global row numbers
// Explicit zeros inside the bands are kept, like the reference's _banded_nzval.
#include <cub/device/device_scan.cuh>

#include "bbm_internal.h"

namespace {

struct SpParams {
    const i64 *R, *C;   // device prefix sums
    i64 Nr, Nc, n;
    int l, u, lam, mu, w, P;
};

__device__ __forceinline__ i64 sp_find_block(const i64* __restrict__ C, i64 Nc, i64 c) {
    i64 lo = 0, hi = Nc;   // largest J with C[J] <= c (skips empty block columns)
    while (hi - lo > 1) {
        const i64 mid = (lo + hi) >> 1;
        if (C[mid] <= c) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) bbb_sparse_count_kernel(const SpParams p, i64* __restrict__ counts) {
    const i64 c = (i64)blockIdx.x * 256 + threadIdx.x;
    if (c > p.n) return;
    if (c == p.n) { counts[c] = 0; return; }
    const i64 J = sp_find_block(p.C, p.Nc, c);
    const i64 j = c - p.C[J];
    i64 cnt = 0;
    if (p.P > 0)
        for (i64 K = max(J - p.u, (i64)0); K <= min(J + p.l, p.Nr - 1); ++K) {
            const i64 rK = p.R[K + 1] - p.R[K];
            const i64 lo = max((i64)0, j - p.mu), hi = min(rK - 1, j + p.lam);
            if (hi >= lo) cnt += hi - lo + 1;
        }
    counts[c] = cnt;
}

template <typename T>
__global__ void __launch_bounds__(256) bbb_sparse_fill_kernel(const SpParams p, const T* __restrict__ data, const i64* __restrict__ colptr, i64 base,
                                                             i64* __restrict__ rowval, T* __restrict__ nzval) {
    const i64 c = (i64)blockIdx.x * 256 + threadIdx.x;
    if (c >= p.n || p.P == 0) return;
    const i64 J = sp_find_block(p.C, p.Nc, c);
    const i64 j = c - p.C[J];
    i64 o = colptr[c] - base;
    const T* col = data + (i64)p.P * c;
    for (i64 K = max(J - p.u, (i64)0); K <= min(J + p.l, p.Nr - 1); ++K) {
        const i64 r0 = p.R[K], rK = p.R[K + 1] - r0;
        const i64 lo = max((i64)0, j - p.mu), hi = min(rK - 1, j + p.lam);
        const T* slab = col + (p.u + K - J) * p.w + p.mu - j;
        for (i64 k = lo; k <= hi; ++k, ++o) {
            rowval[o] = r0 + k + base;
            nzval[o] = slab[k];
        }
    }
}

__global__ void __launch_bounds__(256) sp_add_base_kernel(i64* __restrict__ v, i64 n, i64 base) {
    const i64 i = (i64)blockIdx.x * 256 + threadIdx.x;
    if (i < n) v[i] += base;
}

bool sp_params(bbm_handle_t h, bbm_bbb_desc_t d, SpParams* p) {
    BbbDescView v;
    if (!bbm_bbb_desc_view(d, &v) || v.h != h) return false;
    *p = SpParams{v.dR, v.dC, v.Nr, v.Nc, v.n, (int)v.l, (int)v.u, (int)v.lam, (int)v.mu, (int)(v.lam + v.mu + 1), (int)v.P};
    return true;
}

template <typename T>
int32_t sparse_fill_impl(bbm_handle_t h, bbm_bbb_desc_t d, const T* d_data, const i64* d_colptr, int32_t index_base, i64* d_rowval, T* d_nzval) {
    BBM_REQUIRE_HANDLE(h);
    SpParams p;
    if (!sp_params(h, d, &p)) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BandedBlockBandedMatrix descriptor");
    if (index_base != 0 && index_base != 1) return bbm_fail(h, BBM_ERR_ARG, "index_base must be 0 or 1");
    if (p.n == 0 || p.P == 0) return BBM_OK;
    if (!d_data || !d_colptr || !d_rowval || !d_nzval) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    BBM_CUDA(h, cudaSetDevice(h->device));
    bbb_sparse_fill_kernel<T><<<(unsigned)((p.n + 255) / 256), 256, 0, h->stream>>>(p, d_data, d_colptr, index_base, d_rowval, d_nzval);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "bbb_sparse_fill_kernel launch: %s", cudaGetErrorString(e));
    h->launches += 1;
    return BBM_OK;
}

}  // namespace

extern "C" {

int32_t bbm_bbb_sparse_colptr(bbm_handle_t h, bbm_bbb_desc_t d, int32_t index_base, int64_t* d_colptr, int64_t* nnz) {
    BBM_REQUIRE_HANDLE(h);
    SpParams p;
    if (!sp_params(h, d, &p)) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BandedBlockBandedMatrix descriptor");
    if (index_base != 0 && index_base != 1) return bbm_fail(h, BBM_ERR_ARG, "index_base must be 0 or 1");
    if (!d_colptr) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    BBM_CUDA(h, cudaSetDevice(h->device));
    const i64 len = p.n + 1;
    // counts and the scan's temporary storage live in the handle's workspace
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const i64*)nullptr, (i64*)nullptr, (int)len, h->stream);
    const size_t cbytes = ((size_t)len * sizeof(i64) + 255) / 256 * 256;
    if (len > (i64)0x7fffffff) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "sparse: more than 2^31 columns");
    int32_t rc = bbm_ws_reserve(h, &h->ws_x, &h->ws_x_bytes, cbytes + tmp_bytes + 256);
    if (rc != BBM_OK) return rc;
    i64* counts = static_cast<i64*>(h->ws_x);
    void* tmp = static_cast<char*>(h->ws_x) + cbytes;
    bbb_sparse_count_kernel<<<(unsigned)((len + 255) / 256), 256, 0, h->stream>>>(p, counts);
    BBM_CUDA(h, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, d_colptr, (int)len, h->stream));
    if (index_base) sp_add_base_kernel<<<(unsigned)((len + 255) / 256), 256, 0, h->stream>>>(d_colptr, len, index_base);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "sparse colptr kernels: %s", cudaGetErrorString(e));
    h->launches += 2 + (index_base ? 1 : 0);
    if (nnz) {
        i64 last = 0;
        BBM_CUDA(h, cudaMemcpyAsync(&last, d_colptr + p.n, sizeof(i64), cudaMemcpyDeviceToHost, h->stream));
        BBM_CUDA(h, cudaStreamSynchronize(h->stream));
        *nnz = last - index_base;
    }
    return BBM_OK;
}

int32_t bbm_bbb_sparse_fill_f64(bbm_handle_t h, bbm_bbb_desc_t d, const double* d_data, const int64_t* d_colptr, int32_t index_base,
                                int64_t* d_rowval, double* d_nzval) {
    return sparse_fill_impl<double>(h, d, d_data, d_colptr, index_base, d_rowval, d_nzval);
}
int32_t bbm_bbb_sparse_fill_f32(bbm_handle_t h, bbm_bbb_desc_t d, const float* d_data, const int64_t* d_colptr, int32_t index_base,
                                int64_t* d_rowval, float* d_nzval) {
    return sparse_fill_impl<float>(h, d, d_data, d_colptr, index_base, d_rowval, d_nzval);
}

}  // extern "C"
