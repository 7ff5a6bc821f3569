// Internal declarations shared by the translation units of libbbm_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/bbm_b200.h"

using i64 = int64_t;

// Fused halo push of the multi-GPU band-walk matvec (csrc/dist.cu sets it up, csrc/bbb.cu consumes it).
// Every rank owns a small device block [BbmPeerSig | left halo x2 | right halo x2] that is mapped into both
// neighbours (CUDA IPC).  An extra CTA 0 of every rank's kernel stores the rank's boundary x blocks
// straight into the neighbours' halo buffers of the epoch's parity over NVLink and raises the epoch flag
// there; CTAs whose steps read a halo block column wait for the flag (acquire.sys) and take that step's x
// from the halo buffer.  Buffers alternate by epoch parity, so a push only has to wait for the
// acknowledgement of two epochs ago and never sits on the critical path.
struct BbmPeerSig {
    unsigned data_flag[2];     // [0]: left neighbour filled my left halo for epoch e, [1]: right
    unsigned ack_flag[2];      // [0]: left neighbour consumed what I pushed to it in epoch e, [1]: right
    unsigned count[2];         // local: finished consumer CTAs of my left / right halo this epoch
    unsigned err;              // set if a wait timed out (result of that call is invalid)
    unsigned pad;
};

struct BbmPeerParams {
    int enabled = 0;
    unsigned epoch = 0;
    void* dst[2] = {nullptr, nullptr};            // neighbour halo buffers of this parity (left, right neighbour)
    long long src_off[2] = {0, 0};                // element offsets into my x_local of what I push
    long long cnt[2] = {0, 0};                    // elements to push
    BbmPeerSig* remote[2] = {nullptr, nullptr};   // neighbours' signal blocks
    BbmPeerSig* local = nullptr;
    const void* halo[2] = {nullptr, nullptr};     // my halo buffers of this parity
    long long halo_col0[2] = {0, 0};              // first local column each halo buffer stands for
    int own_J0 = 0, own_J1 = 0;                   // local block-column range this rank owns
    int has_halo[2] = {0, 0};                     // a neighbour fills my left / right halo
    // Acknowledgements are only needed across ONE-SIDED neighbour pairs.  Where halos flow both ways, the
    // neighbour's kernel e cannot start before its kernel e-1 completed, which needed my push of epoch e-1,
    // whose kernel started after my kernel e-2 completed -- so the parity buffer epoch e overwrites was last
    // read by a finished kernel, without any explicit signal (and without remote stores in the kernel tail).
    int ack_wait[2] = {0, 0};                     // I push to that side but receive nothing from it
    int ack_send[2] = {0, 0};                     // I receive from that side but push nothing to it
    int consumers[2] = {0, 0};                    // CTAs reading each halo (filled by the launcher)
    long long timeout_clk = 0;                    // SM clocks a peer wait may take; 0 = wait for ever
    unsigned* err_host = nullptr;                 // host-mapped word latched by a timed-out wait (checked by the next call)
};

struct bbm_context {
    uint32_t magic = 0xBB200C0Du;
    int device = 0;
    cudaStream_t stream = nullptr;
    int num_sms = 148;
    size_t smem_optin = 0;
    std::string err;
    std::map<std::string, i64> opts;
    i64 opt_gen = 0;   // bumped by bbm_set_option: cached launch plans are keyed on it
    i64 launches = 0;
    double last_cond = 0.0;   // largest ||T_KK|| ||inv(T_KK)|| seen by the last block triangular solve's condition guard
    // workspace for the *_host_* calls (device copies of X and Y)
    void* ws_x = nullptr; size_t ws_x_bytes = 0;
    void* ws_y = nullptr; size_t ws_y_bytes = 0;
    void* ws_z = nullptr; size_t ws_z_bytes = 0;   // triangular solve: inverses of the diagonal blocks
    void* ws_w = nullptr; size_t ws_w_bytes = 0;   // triangular solve: level temporaries / wave-ordered coefficients
    // low-priority side stream (L2 prefetch ahead of the triangular solve's sequential chain)
    const BbmPeerParams* peer = nullptr;  // non-null only inside a fused multi-GPU product call
    cudaStream_t copy_in = nullptr, copy_out = nullptr;  // pipelined host-vector product
    cudaStream_t side_stream = nullptr;
    cudaEvent_t side_ev[2] = {nullptr, nullptr};
    cudaEvent_t qr_ev[4] = {nullptr, nullptr, nullptr, nullptr};   // qr! look-ahead: panel done x2, rest-of-update done x2
};

inline bool bbm_valid(bbm_handle_t h) { return h != nullptr && h->magic == 0xBB200C0Du; }

inline int32_t bbm_fail(bbm_handle_t h, int32_t code, const char* fmt, ...) {
    if (bbm_valid(h)) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        h->err = buf;
    }
    return code;
}

#define BBM_CUDA(h, expr)                                                                         \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return bbm_fail((h), BBM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                                  \
    } while (0)

#define BBM_REQUIRE_HANDLE(h) \
    do { if (!bbm_valid(h)) return BBM_ERR_HANDLE; } while (0)

inline i64 bbm_opt(bbm_handle_t h, const char* key, i64 dflt) {
    auto it = h->opts.find(key);
    return it == h->opts.end() ? dflt : it->second;
}

// grow-only device workspace
int32_t bbm_ws_reserve(bbm_handle_t h, void** ws, size_t* have, size_t need);

// read-only view of a BandedBlockBandedMatrix descriptor for the other translation units
struct BbbDescView {
    bbm_handle_t h;
    i64 Nr, Nc, l, u, lam, mu, m, n, P, maxr;
    const i64 *r, *c;    // host block sizes
    const i64 *dR, *dC;  // device prefix sums (Nr+1, Nc+1)
};
bool bbm_bbb_desc_view(bbm_bbb_desc_t d, BbbDescView* out);

// read-only view of a BlockSkylineMatrix descriptor (host tables + device copies) for the other translation units
struct SkyDescView {
    bbm_handle_t h;
    i64 Nr, Nc, m, n, numel;
    const i64 *r, *c, *R, *C, *klo, *khi, *off, *S;           // host
    const i64 *d_off, *d_S, *d_klo, *d_khi, *d_R, *d_C;       // device
};
bool bbm_sky_desc_view(bbm_sky_desc_t d, SkyDescView* out);

// Pieces of the skyline products for the sharded paths (csrc/distsky.cu).  elem = sizeof(T).
//  bbm_sky_mul_rows:     only block rows [Klo,Khi); phase 0 self-contained, 1 first piece (pre-scales all of y when the
//                        split kernel is in use), 2 later piece
//  bbm_sky_matmul_phase: phase 0 whole product, 1 tiles that only multiply B's owned block rows [own_lo,own_hi), 2 the rest
int32_t bbm_sky_mul_rows(bbm_handle_t h, bbm_sky_desc_t d, int elem, double alpha, const void* data, const void* X, i64 ldx, i64 nrhs, double beta,
                         void* Y, i64 ldy, i64 Klo, i64 Khi, int phase);
int32_t bbm_sky_matmul_phase(bbm_handle_t h, bbm_sky_desc_t A, bbm_sky_desc_t B, bbm_sky_desc_t C, int elem, double alpha, const void* dA,
                             const void* dB, double beta, void* dC, int phase, i64 own_lo, i64 own_hi);
