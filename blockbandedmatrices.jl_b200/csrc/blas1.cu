// Level-1 operations on the storage arrays of same-structure block-banded matrices (and on vectors):
//   axpy!(a, A, B)  for BandedBlockBandedMatrix / BlockSkylineMatrix operands with identical structure is
//                   axpy! on `data` (src/BandedBlockBandedMatrix.jl:605-606, src/broadcast.jl:307-314);
//   lmul!(a, A) / rmul!(A, a) scale `data` (src/BandedBlockBandedMatrix.jl:609-612, src/broadcast.jl:219-252).
// Unused storage slots are scaled / added like the rest -- they are never read as matrix entries.
// HBM-bound streaming kernels: 16-byte vector accesses, grid = a few waves of 148 SMs.
#include <algorithm>

#include "bbm_internal.h"

namespace {

template <typename T, typename V, int VEC>
__global__ void __launch_bounds__(256) axpy_kernel(i64 n, T alpha, const T* __restrict__ x, T* __restrict__ y) {
    const i64 nv = n / VEC;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const V* xv = reinterpret_cast<const V*>(x);
    V* yv = reinterpret_cast<V*>(y);
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        V a = xv[i], b = yv[i];
        T* pa = reinterpret_cast<T*>(&a);
        T* pb = reinterpret_cast<T*>(&b);
#pragma unroll
        for (int k = 0; k < VEC; ++k) pb[k] = fma(alpha, pa[k], pb[k]);
        yv[i] = b;
    }
    for (i64 i = nv * VEC + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = fma(alpha, x[i], y[i]);
}

template <typename T, typename V, int VEC>
__global__ void __launch_bounds__(256) scal_kernel(i64 n, T alpha, T* __restrict__ x) {
    const i64 nv = n / VEC;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    V* xv = reinterpret_cast<V*>(x);
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        V a = xv[i];
        T* pa = reinterpret_cast<T*>(&a);
#pragma unroll
        for (int k = 0; k < VEC; ++k) pa[k] = alpha * pa[k];
        xv[i] = a;
    }
    for (i64 i = nv * VEC + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = alpha * x[i];
}

template <typename T>
int32_t axpy_impl(bbm_handle_t h, i64 n, T alpha, const T* x, T* y) {
    BBM_REQUIRE_HANDLE(h);
    if (n < 0) return bbm_fail(h, BBM_ERR_ARG, "n < 0");
    if (n == 0) return BBM_OK;
    if (!x || !y) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    BBM_CUDA(h, cudaSetDevice(h->device));
    const unsigned grid = (unsigned)std::min<i64>((n + 1023) / 1024, (i64)h->num_sms * 16);
    const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    if (al && sizeof(T) == 8) axpy_kernel<T, double2, 2><<<grid, 256, 0, h->stream>>>(n, alpha, x, y);
    else if (al) axpy_kernel<T, float4, 4><<<grid, 256, 0, h->stream>>>(n, alpha, x, y);
    else axpy_kernel<T, T, 1><<<grid, 256, 0, h->stream>>>(n, alpha, x, y);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "axpy launch: %s", cudaGetErrorString(e));
    h->launches += 1;
    return BBM_OK;
}

template <typename T>
int32_t scal_impl(bbm_handle_t h, i64 n, T alpha, T* x) {
    BBM_REQUIRE_HANDLE(h);
    if (n < 0) return bbm_fail(h, BBM_ERR_ARG, "n < 0");
    if (n == 0) return BBM_OK;
    if (!x) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    BBM_CUDA(h, cudaSetDevice(h->device));
    const unsigned grid = (unsigned)std::min<i64>((n + 1023) / 1024, (i64)h->num_sms * 16);
    const bool al = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    if (al && sizeof(T) == 8) scal_kernel<T, double2, 2><<<grid, 256, 0, h->stream>>>(n, alpha, x);
    else if (al) scal_kernel<T, float4, 4><<<grid, 256, 0, h->stream>>>(n, alpha, x);
    else scal_kernel<T, T, 1><<<grid, 256, 0, h->stream>>>(n, alpha, x);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "scal launch: %s", cudaGetErrorString(e));
    h->launches += 1;
    return BBM_OK;
}

}  // namespace

extern "C" {
int32_t bbm_axpy_f64(bbm_handle_t h, int64_t n, double alpha, const double* d_x, double* d_y) { return axpy_impl<double>(h, n, alpha, d_x, d_y); }
int32_t bbm_axpy_f32(bbm_handle_t h, int64_t n, float alpha, const float* d_x, float* d_y) { return axpy_impl<float>(h, n, alpha, d_x, d_y); }
int32_t bbm_scal_f64(bbm_handle_t h, int64_t n, double alpha, double* d_x) { return scal_impl<double>(h, n, alpha, d_x); }
int32_t bbm_scal_f32(bbm_handle_t h, int64_t n, float alpha, float* d_x) { return scal_impl<float>(h, n, alpha, d_x); }
}
