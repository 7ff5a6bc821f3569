// Handle, option, and memory-utility entry points of the C ABI (include/bbm_b200.h).
#include <cstring>

#include "bbm_internal.h"

static const char* const kOptionKeys[] = {
    "bbb.tile", "bbb.stages", "bbb.items_per_sm", "bbb.kernel", "bbb.maxseg", "bbb.l2hint", "bbb.l2keep_pct", "bbb.waves", "bbb.host_chunks", "bbb.ovh", "bbb.mm_kernel", "bbb.mm_staged", "bbb.pdl", "bbb.ctas_per_sm",
    "sky.kernel", "sky.mv_split", "sky.mm_kernel", "sky.mm_bk", "sky.mm_ws", "sky.mm_f32_kernel", "bbb.mm_ctas_per_sm", "qr.panel", "qr.flow", "qr.lookahead", "trsm.kernel", "trsm.prefetch", "trsm.pdl", "trsm.cond_check", "trsm.cond_max", "trsm.flow_ctas_per_sm", "trsm.flow_bk", "trsm.flow_ctr_stride", "trsm.flow_sleep_ns", "trsm.last_cond_x1000", "trsm.last_path", "tri.pf", "tri.sync", "tri.kernel", "tri.lean", "tri.push", "tri.ws_max_mb", "tri.rhs_batch", "dist.graph", "dist.peer", "dist.lead_steps", "dist.peer_timeout_ms", nullptr};

extern "C" {

int32_t bbm_version(void) { return 100; }

const char* bbm_status_string(int32_t s) {
    switch (s) {
        case BBM_OK: return "ok";
        case BBM_ERR_HANDLE: return "invalid handle or descriptor";
        case BBM_ERR_DIM: return "DimensionMismatch";
        case BBM_ERR_BAND: return "BandError";
        case BBM_ERR_ARG: return "invalid argument";
        case BBM_ERR_CUDA: return "CUDA error";
        case BBM_ERR_NCCL: return "peer/collective error";
        case BBM_ERR_ALLOC: return "allocation failed";
        case BBM_ERR_UNSUPPORTED: return "not supported by the device path";
        default: return "unknown status";
    }
}

int32_t bbm_create(int32_t device, bbm_handle_t* out) {
    if (!out) return BBM_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return BBM_ERR_CUDA;  // no CPU fallback
    if (device < 0 || device >= count) return BBM_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return BBM_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return BBM_ERR_CUDA;
    if (prop.major < 10) return BBM_ERR_UNSUPPORTED;  // kernels are built for sm_100a only
    bbm_context* h = new (std::nothrow) bbm_context();
    if (!h) return BBM_ERR_ALLOC;
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    *out = h;
    return BBM_OK;
}

int32_t bbm_destroy(bbm_handle_t h) {
    BBM_REQUIRE_HANDLE(h);
    cudaSetDevice(h->device);
    if (h->ws_x) cudaFree(h->ws_x);
    if (h->ws_y) cudaFree(h->ws_y);
    if (h->ws_z) cudaFree(h->ws_z);
    if (h->ws_w) cudaFree(h->ws_w);
    if (h->copy_in) cudaStreamDestroy(h->copy_in);
    if (h->copy_out) cudaStreamDestroy(h->copy_out);
    if (h->side_stream) { cudaStreamSynchronize(h->side_stream); cudaStreamDestroy(h->side_stream); }
    for (int i = 0; i < 2; ++i) if (h->side_ev[i]) cudaEventDestroy(h->side_ev[i]);
    for (int i = 0; i < 4; ++i) if (h->qr_ev[i]) cudaEventDestroy(h->qr_ev[i]);
    h->magic = 0;
    delete h;
    return BBM_OK;
}

int32_t bbm_set_stream(bbm_handle_t h, void* s) {
    BBM_REQUIRE_HANDLE(h);
    h->stream = reinterpret_cast<cudaStream_t>(s);
    return BBM_OK;
}

int32_t bbm_synchronize(bbm_handle_t h) {
    BBM_REQUIRE_HANDLE(h);
    BBM_CUDA(h, cudaSetDevice(h->device));
    BBM_CUDA(h, cudaStreamSynchronize(h->stream));
    return BBM_OK;
}

const char* bbm_last_error(bbm_handle_t h) { return bbm_valid(h) ? h->err.c_str() : "invalid handle"; }

int32_t bbm_set_option(bbm_handle_t h, const char* key, int64_t value) {
    BBM_REQUIRE_HANDLE(h);
    if (!key) return bbm_fail(h, BBM_ERR_ARG, "null option key");
    for (int i = 0; kOptionKeys[i]; ++i)
        if (strcmp(kOptionKeys[i], key) == 0) { h->opts[key] = value; ++h->opt_gen; return BBM_OK; }
    return bbm_fail(h, BBM_ERR_ARG, "unknown option '%s'", key);
}

int32_t bbm_get_option(bbm_handle_t h, const char* key, int64_t* value) {
    BBM_REQUIRE_HANDLE(h);
    if (!key || !value) return bbm_fail(h, BBM_ERR_ARG, "null argument");
    for (int i = 0; kOptionKeys[i]; ++i)
        if (strcmp(kOptionKeys[i], key) == 0) { *value = bbm_opt(h, key, 0); return BBM_OK; }
    return bbm_fail(h, BBM_ERR_ARG, "unknown option '%s'", key);
}

int32_t bbm_launch_count(bbm_handle_t h, int64_t* count) {
    BBM_REQUIRE_HANDLE(h);
    if (!count) return bbm_fail(h, BBM_ERR_ARG, "null argument");
    *count = h->launches;
    return BBM_OK;
}

int32_t bbm_malloc(bbm_handle_t h, void** d_ptr, size_t bytes) {
    BBM_REQUIRE_HANDLE(h);
    if (!d_ptr) return bbm_fail(h, BBM_ERR_ARG, "null argument");
    *d_ptr = nullptr;
    BBM_CUDA(h, cudaSetDevice(h->device));
    // The band-walk kernel moves `data` and `x` with 16-byte-granular bulk copies: it may read up to 15 bytes past the last
    // element it addresses.  Every bbm_malloc allocation therefore ends with 32 bytes of readable padding of its own (cudaMalloc's
    // 256-byte granularity already guarantees it today; a sub-allocating pool behind this call would not).
    bytes = ((bytes + 15) & ~size_t(15)) + 32;
    cudaError_t e = cudaMalloc(d_ptr, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return bbm_fail(h, BBM_ERR_ALLOC, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); }
    return BBM_OK;
}

int32_t bbm_free(bbm_handle_t h, void* d_ptr) {
    BBM_REQUIRE_HANDLE(h);
    BBM_CUDA(h, cudaSetDevice(h->device));
    BBM_CUDA(h, cudaFree(d_ptr));
    return BBM_OK;
}

int32_t bbm_host_alloc(bbm_handle_t h, void** h_ptr, size_t bytes) {
    BBM_REQUIRE_HANDLE(h);
    if (!h_ptr) return bbm_fail(h, BBM_ERR_ARG, "null argument");
    *h_ptr = nullptr;
    BBM_CUDA(h, cudaSetDevice(h->device));
    cudaError_t e = cudaMallocHost(h_ptr, bytes ? bytes : 16);
    if (e != cudaSuccess) { cudaGetLastError(); return bbm_fail(h, BBM_ERR_ALLOC, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e)); }
    return BBM_OK;
}

int32_t bbm_host_free(bbm_handle_t h, void* h_ptr) {
    BBM_REQUIRE_HANDLE(h);
    BBM_CUDA(h, cudaSetDevice(h->device));
    BBM_CUDA(h, cudaFreeHost(h_ptr));
    return BBM_OK;
}

static int32_t copy_async(bbm_handle_t h, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
    BBM_REQUIRE_HANDLE(h);
    if (bytes == 0) return BBM_OK;
    if (!dst || !src) return bbm_fail(h, BBM_ERR_ARG, "null pointer in memcpy");
    BBM_CUDA(h, cudaSetDevice(h->device));
    BBM_CUDA(h, cudaMemcpyAsync(dst, src, bytes, kind, h->stream));
    return BBM_OK;
}

int32_t bbm_memcpy_h2d(bbm_handle_t h, void* d, const void* s, size_t n) { return copy_async(h, d, s, n, cudaMemcpyHostToDevice); }
int32_t bbm_memcpy_d2h(bbm_handle_t h, void* d, const void* s, size_t n) { return copy_async(h, d, s, n, cudaMemcpyDeviceToHost); }
int32_t bbm_memcpy_d2d(bbm_handle_t h, void* d, const void* s, size_t n) { return copy_async(h, d, s, n, cudaMemcpyDeviceToDevice); }

int32_t bbm_memset(bbm_handle_t h, void* d, int32_t v, size_t n) {
    BBM_REQUIRE_HANDLE(h);
    if (n == 0) return BBM_OK;
    if (!d) return bbm_fail(h, BBM_ERR_ARG, "null pointer in memset");
    BBM_CUDA(h, cudaSetDevice(h->device));
    BBM_CUDA(h, cudaMemsetAsync(d, v, n, h->stream));
    return BBM_OK;
}

}  // extern "C"

int32_t bbm_ws_reserve(bbm_handle_t h, void** ws, size_t* have, size_t need) {
    if (*have >= need) return BBM_OK;
    BBM_CUDA(h, cudaStreamSynchronize(h->stream));
    if (*ws) { BBM_CUDA(h, cudaFree(*ws)); *ws = nullptr; *have = 0; }
    size_t bytes = need + need / 8 + 256;
    cudaError_t e = cudaMalloc(ws, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return bbm_fail(h, BBM_ERR_ALLOC, "workspace cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); }
    *have = bytes;
    return BBM_OK;
}
