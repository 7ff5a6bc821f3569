// Row-block-partitioned BandedBlockBandedMatrix matvec over several B200s (one process per GPU).
//
// Reference precedent: the row-block data parallelism of examples/shared_array_backend.jl:92-122
// (balanced contiguous block-row ranges, read-only shared x).  Here x is distributed too, so the
// x block-columns inside the l+u block halo are exchanged: NCCL send/recv between the owners,
// overlapped with the product of the interior block rows, which need no halo.
//
// A contiguous block-row range of a BandedBlockBandedMatrix is itself banded-block-banded with
// shifted block bandwidths (l-s, u+s) over a column range of the same band storage
// (src/BandedBlockBandedMatrix.jl:450-456 sub-view bandwidths, :416-420 sub-view data), so every
// rank -- and the interior / boundary pieces inside a rank -- runs the single-GPU band-walk kernel
// on a column slab of the parent `data`, unchanged.
//
// The layout planning (bbm_bbb_partition_rows / bbm_bbb_dist_layout) is pure host code: it needs
// no GPU and is what the world_size-2 gloo tests exercise on CPU.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <new>

#include "bbm_internal.h"
#include "dist_internal.h"

BbmNccl& bbm_nccl() {
    static BbmNccl n;
    return n;
}

bool bbm_nccl_load() {
    BbmNccl& n = bbm_nccl();
    if (n.lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    void* lib = nullptr;
    for (int i = 0; names[i] && !lib; ++i) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { n.err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
#define BIND(field, sym)                                             \
    n.field = reinterpret_cast<decltype(n.field)>(dlsym(lib, sym)); \
    if (!n.field) { n.err = std::string("missing NCCL symbol ") + sym; return false; }
    BIND(GetUniqueId, "ncclGetUniqueId")
    BIND(CommInitRank, "ncclCommInitRank")
    BIND(CommDestroy, "ncclCommDestroy")
    BIND(Send, "ncclSend")
    BIND(Recv, "ncclRecv")
    BIND(AllGather, "ncclAllGather")
    BIND(GroupStart, "ncclGroupStart")
    BIND(GroupEnd, "ncclGroupEnd")
    BIND(GetErrorString, "ncclGetErrorString")
    BIND(GetVersion, "ncclGetVersion")
#undef BIND
    n.lib = lib;
    return true;
}

namespace {

inline BbmNccl& nccl() { return bbm_nccl(); }
inline bool nccl_load() { return bbm_nccl_load(); }
typedef BbmNccl Nccl;

inline i64 clampi(i64 v, i64 lo, i64 hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct RankRanges {
    i64 K0, K1;      // owned block rows
    i64 J0, J1;      // owned block columns (x ownership)
    i64 Nlo, Nhi;    // block columns the owned rows touch ("needed"), empty if K0 == K1
};

RankRanges rank_ranges(i64 Nr, i64 Nc, i64 l, i64 u, int nranks, int p, const i64* Kb) {
    RankRanges g;
    g.K0 = Kb[p]; g.K1 = Kb[p + 1];
    g.J0 = std::min(g.K0, Nc);
    g.J1 = (p == nranks - 1) ? Nc : std::min(g.K1, Nc);
    if (g.J1 < g.J0) g.J1 = g.J0;
    if (g.K1 > g.K0 && l + u + 1 > 0) {
        g.Nlo = clampi(g.K0 - l, 0, Nc);
        g.Nhi = clampi(g.K1 - 1 + u + 1, 0, Nc);
        if (g.Nhi < g.Nlo) g.Nhi = g.Nlo;
    } else {
        g.Nlo = g.Nhi = g.J0;
    }
    return g;
}

}  // namespace

struct GraphKey {
    const void* data = nullptr; void* x = nullptr; void* y = nullptr;
    double alpha = 0, beta = 0; i64 nrhs = 0; int ts = 0; cudaStream_t stream = nullptr;
    bool operator==(const GraphKey& o) const {
        return data == o.data && x == o.x && y == o.y && alpha == o.alpha && beta == o.beta && nrhs == o.nrhs &&
               ts == o.ts && stream == o.stream;
    }
};

struct bbm_dist_bbb_plan {
    uint32_t magic = 0xD157B200u;
    // global structure (neighbour layouts are recomputed from it)
    i64 Nr = 0, Nc = 0, l = 0, u = 0, lam = 0, mu = 0;
    std::vector<i64> r, c, kb;
    // fused halo push over NVLink peer memory (bbm_dist_bbb_register_x)
    bool peer_ok = false;
    int reg_elem = 0;
    unsigned epoch = 0;
    BbmPeerParams peer_tpl;
    BbmPeerSig* sig = nullptr;
    void* ipc_sig[2] = {nullptr, nullptr};   // neighbours' blocks mapped here
    size_t own_buf[2][2] = {{0, 0}, {0, 0}}; // [side][parity] byte offsets of my halo buffers in my block
    size_t nb_buf[2][2] = {{0, 0}, {0, 0}};  // [side][parity] byte offsets of the buffers I fill in the neighbours' blocks
    bbm_bbb_desc_t desc_all = nullptr;
    unsigned* err_host = nullptr;     // host-mapped word a timed-out peer wait latches (checked before every product)
    unsigned* err_host_dev = nullptr; // its device address
    int clock_khz = 0;                // SM clock (peer-wait time-out in clocks)
    cudaGraphExec_t gexec = nullptr;  // cached CUDA graph of one whole step
    GraphKey gkey;
    i64 glaunches = 0;
    bool gfailed = false;
    bbm_comm* comm = nullptr;
    bbm_handle_t h = nullptr;
    bbm_dist_layout lay{};
    i64 P = 0;
    std::vector<i64> sends, recvs;  // triples (peer, local column offset, count)
    // sub-problems: 0 interior, 1 left boundary, 2 right boundary; each (desc, data col offset, y row offset)
    bbm_bbb_desc_t desc[3] = {nullptr, nullptr, nullptr};
    i64 coloff[3] = {0, 0, 0}, rowoff[3] = {0, 0, 0};
    // pipelined host-vector product (bbm_dist_bbb_mul_host_*): block-row chunks of the owned rows as sub-views of the
    // local slab, device workspaces for x_local / y_own / the staged send blocks, events
    struct HostChunk { i64 Ka, Kb, Jl, Jh; bool halo; bbm_bbb_desc_t sub; };
    std::vector<HostChunk> hchunks;
    std::vector<cudaEvent_t> hev;
    std::vector<i64> C_local, R_own;   // prefix sums of the local block columns / owned block rows (elements)
    void* ws_x = nullptr; void* ws_y = nullptr; void* ws_s = nullptr;
    int ws_elem = 0;
};

namespace {
inline bool comm_valid(bbm_comm* c) { return bbm_comm_valid(c); }
inline bool plan_valid(bbm_dist_bbb_plan* p) { return p && p->magic == 0xD157B200u && bbm_valid(p->h); }

int32_t layout_impl(i64 Nr, i64 Nc, const i64* r, const i64* c, i64 l, i64 u, int nranks, int rank, const i64* Kb,
                    bbm_dist_layout* out, std::vector<i64>* sends, std::vector<i64>* recvs) {
    if (Nr < 0 || Nc < 0 || nranks < 1 || rank < 0 || rank >= nranks || !Kb || !out) return BBM_ERR_ARG;
    if (Kb[0] != 0 || Kb[nranks] != Nr) return BBM_ERR_ARG;
    for (int p = 0; p < nranks; ++p)
        if (Kb[p + 1] < Kb[p]) return BBM_ERR_ARG;
    std::vector<i64> R(Nr + 1, 0), C(Nc + 1, 0);
    for (i64 K = 0; K < Nr; ++K) R[K + 1] = R[K] + r[K];
    for (i64 J = 0; J < Nc; ++J) C[J + 1] = C[J] + c[J];
    const RankRanges me = rank_ranges(Nr, Nc, l, u, nranks, rank, Kb);
    bbm_dist_layout L{};
    L.K0 = me.K0; L.K1 = me.K1; L.J0 = me.J0; L.J1 = me.J1;
    L.Jlo = std::min(me.J0, me.Nlo);
    L.Jhi = std::max(me.J1, me.Nhi);
    if (me.Nlo == me.Nhi) { L.Jlo = me.J0; L.Jhi = me.J1; }
    L.m_own = R[me.K1] - R[me.K0];
    L.n_own = C[me.J1] - C[me.J0];
    L.n_local = C[L.Jhi] - C[L.Jlo];
    L.col_offset = C[L.Jlo];
    L.own_offset = C[me.J0] - C[L.Jlo];
    L.row_offset = R[me.K0];
    // interior block rows: every touched block column is owned by this rank
    i64 Ka = me.K0, Kbnd = me.K1;
    if (me.K1 > me.K0 && l + u + 1 > 0) {
        if (me.Nlo < me.J0) Ka = clampi(me.J0 + l, me.K0, me.K1);
        if (me.Nhi > me.J1) Kbnd = clampi(me.J1 - u, Ka, me.K1);
        // rows whose support leaves the owned range on the far side (only with very negative bandwidths)
        if (me.Nlo >= me.J1 || me.Nhi <= me.J0) { Ka = me.K0; Kbnd = me.K0; }
    }
    L.Ka = Ka; L.Kb = Kbnd;
    // exchange lists: what I need from q / what q needs from me, as local-column ranges
    if (sends) sends->clear();
    if (recvs) recvs->clear();
    i64 ns = 0, nr_ = 0;
    for (int q = 0; q < nranks; ++q) {
        if (q == rank) continue;
        const RankRanges o = rank_ranges(Nr, Nc, l, u, nranks, q, Kb);
        const i64 rlo = std::max(me.Nlo, o.J0), rhi = std::min(me.Nhi, o.J1);  // I need, q owns
        if (rhi > rlo && C[rhi] > C[rlo]) {
            ++nr_;
            if (recvs) { recvs->push_back(q); recvs->push_back(C[rlo] - C[L.Jlo]); recvs->push_back(C[rhi] - C[rlo]); }
        }
        const i64 slo = std::max(o.Nlo, me.J0), shi = std::min(o.Nhi, me.J1);  // q needs, I own
        if (shi > slo && C[shi] > C[slo]) {
            ++ns;
            if (sends) { sends->push_back(q); sends->push_back(C[slo] - C[L.Jlo]); sends->push_back(C[shi] - C[slo]); }
        }
    }
    L.nsend = ns; L.nrecv = nr_;
    *out = L;
    return BBM_OK;
}

// Enqueue one distributed product.  Stream plan:
//   compute stream : [x ready] ------------------ interior rows ------------------ [join]
//   exchange stream:      wait -> NCCL send/recv -> boundary rows (high priority) --^
// The boundary pieces write disjoint rows of y, so they run concurrently with the interior kernel
// and are absorbed into its tail instead of adding two serial launches per step.
template <typename T>
int32_t dist_enqueue(bbm_dist_bbb_plan* pl, T alpha, const T* d_data_local, T* d_x_local, i64 nrhs, T beta, T* d_y_own) {
    bbm_handle_t h = pl->h;
    bbm_comm* cm = pl->comm;
    const bbm_dist_layout& L = pl->lay;
    const i64 ldx = std::max<i64>(1, L.n_local), ldy = std::max<i64>(1, L.m_own);
    const bool exchange = cm && cm->comm && cm->nranks > 1 && (!pl->sends.empty() || !pl->recvs.empty());
    auto run = [&](int which) -> int32_t {
        if (!pl->desc[which]) return BBM_OK;
        const T* data = d_data_local + pl->P * pl->coloff[which];
        const T* x = d_x_local + pl->coloff[which];
        T* y = d_y_own + pl->rowoff[which];
        if (sizeof(T) == 8)
            return bbm_bbb_mul_f64(h, pl->desc[which], (double)alpha, (const double*)data, (const double*)x, ldx, nrhs,
                                   (double)beta, (double*)y, ldy);
        return bbm_bbb_mul_f32(h, pl->desc[which], (float)alpha, (const float*)data, (const float*)x, ldx, nrhs,
                               (float)beta, (float*)y, ldy);
    };
    if (!exchange) {
        int32_t rc = run(0);
        if (rc == BBM_OK) rc = run(1);
        if (rc == BBM_OK) rc = run(2);
        return rc;
    }
    Nccl& n = nccl();
    // the exchange may start once everything queued on the compute stream so far (the producer of
    // x, and the previous call's readers of the halo) has finished
    BBM_CUDA(h, cudaEventRecord(cm->ev_ready, h->stream));
    BBM_CUDA(h, cudaStreamWaitEvent(cm->stream, cm->ev_ready, 0));
    ncclResult_t nrc = n.GroupStart();
    for (i64 q = 0; q < nrhs && nrc == ncclSuccess; ++q) {
        T* xq = d_x_local + q * ldx;
        for (size_t i = 0; i + 2 < pl->recvs.size() && nrc == ncclSuccess; i += 3)
            nrc = n.Recv(xq + pl->recvs[i + 1], (size_t)pl->recvs[i + 2] * sizeof(T), ncclChar, (int)pl->recvs[i], cm->comm, cm->stream);
        for (size_t i = 0; i + 2 < pl->sends.size() && nrc == ncclSuccess; i += 3)
            nrc = n.Send(xq + pl->sends[i + 1], (size_t)pl->sends[i + 2] * sizeof(T), ncclChar, (int)pl->sends[i], cm->comm, cm->stream);
    }
    ncclResult_t nrc2 = n.GroupEnd();
    if (nrc == ncclSuccess) nrc = nrc2;
    if (nrc != ncclSuccess) return bbm_fail(h, BBM_ERR_NCCL, "halo exchange: %s", n.GetErrorString(nrc));
    // boundary rows behind the exchange, on its (high-priority) stream
    cudaStream_t compute = h->stream;
    h->stream = cm->stream;
    int32_t rc = run(1);
    if (rc == BBM_OK) rc = run(2);
    h->stream = compute;
    if (rc != BBM_OK) return rc;
    BBM_CUDA(h, cudaEventRecord(cm->ev_halo, cm->stream));
    rc = run(0);  // interior rows: need no halo
    if (rc != BBM_OK) return rc;
    BBM_CUDA(h, cudaStreamWaitEvent(h->stream, cm->ev_halo, 0));
    return BBM_OK;
}

template <typename T>
int32_t dist_mul_impl(bbm_dist_bbb_plan* pl, T alpha, const T* d_data_local, T* d_x_local, i64 nrhs, T beta, T* d_y_own) {
    if (!plan_valid(pl)) return BBM_ERR_HANDLE;
    bbm_handle_t h = pl->h;
    bbm_comm* cm = pl->comm;
    if (nrhs < 0) return bbm_fail(h, BBM_ERR_ARG, "nrhs < 0");
    if (nrhs == 0) return BBM_OK;
    BBM_CUDA(h, cudaSetDevice(h->device));
    const bool exchange = cm && cm->comm && cm->nranks > 1 && (!pl->sends.empty() || !pl->recvs.empty());
    if (pl->peer_ok && nrhs == 1 && (int)sizeof(T) == pl->reg_elem && bbm_opt(h, "dist.peer", 1) != 0) {
        // ONE kernel per step: a pusher CTA stores the boundary x blocks into the neighbours' halo buffers
        // over NVLink peer memory, halo-reading CTAs wait on the epoch flag, everything else streams as on
        // one GPU.  (The halo parts of x_local are neither read nor written on this path.)
        // a peer wait of an earlier product expired: that product's rows were poisoned with NaN, the epochs of the
        // ranks no longer agree -- refuse (every rank sees its own or its neighbour's time-out) instead of computing
        // on stale halos.  bbm_dist_bbb_enable_peer re-arms the path.
        if (pl->err_host && *reinterpret_cast<volatile unsigned*>(pl->err_host) != 0)
            return bbm_fail(h, BBM_ERR_NCCL, "fused halo push: a peer wait timed out in an earlier product (results since are invalid); "
                                             "call bbm_dist_bbb_enable_peer again or set dist.peer=0");
        BbmPeerParams pp = pl->peer_tpl;
        pp.enabled = 1;
        pp.epoch = ++pl->epoch;
        {   // "dist.peer_timeout_ms": how long a wait for a neighbour may take (default 30 s; 0 = for ever, like NCCL)
            const i64 ms = bbm_opt(h, "dist.peer_timeout_ms", 30000);
            pp.timeout_clk = ms <= 0 ? 0 : (long long)ms * (long long)std::max(1, pl->clock_khz);
            pp.err_host = pl->err_host_dev;
        }
        const int par = (int)(pp.epoch & 1u);
        for (int sd = 0; sd < 2; ++sd) {
            pp.halo[sd] = pp.has_halo[sd] ? reinterpret_cast<const char*>(pl->sig) + pl->own_buf[sd][par] : nullptr;
            pp.dst[sd] = (pp.cnt[sd] > 0 && pl->ipc_sig[sd]) ? static_cast<char*>(pl->ipc_sig[sd]) + pl->nb_buf[sd][par] : nullptr;
        }
        const bbm_dist_layout& L = pl->lay;
        h->peer = &pp;
        int32_t rc;
        if (sizeof(T) == 8)
            rc = bbm_bbb_mul_f64(h, pl->desc_all, (double)alpha, (const double*)d_data_local, (const double*)d_x_local,
                                 std::max<i64>(1, L.n_local), 1, (double)beta, (double*)d_y_own, std::max<i64>(1, L.m_own));
        else
            rc = bbm_bbb_mul_f32(h, pl->desc_all, (float)alpha, (const float*)d_data_local, (const float*)d_x_local,
                                 std::max<i64>(1, L.n_local), 1, (float)beta, (float*)d_y_own, std::max<i64>(1, L.m_own));
        h->peer = nullptr;
        return rc;
    }
    const bool want_graph = exchange && bbm_opt(h, "dist.graph", 0) != 0;
    if (!want_graph) return dist_enqueue<T>(pl, alpha, d_data_local, d_x_local, nrhs, beta, d_y_own);

    // Repeated products with the same buffers (every iterative solver) replay one CUDA graph of the
    // whole step -- exchange, boundary and interior kernels, cross-stream edges -- so the host
    // issues one launch per step.  First call with a new argument set runs eagerly (it may allocate
    // the work partition), the second captures, later ones replay.
    GraphKey key{d_data_local, d_x_local, d_y_own, (double)alpha, (double)beta, nrhs, (int)sizeof(T), h->stream};
    if (pl->gexec && key == pl->gkey) {
        BBM_CUDA(h, cudaGraphLaunch(pl->gexec, h->stream));
        h->launches += pl->glaunches;
        return BBM_OK;
    }
    if (!(key == pl->gkey) || pl->gfailed) {
        if (pl->gexec) { cudaGraphExecDestroy(pl->gexec); pl->gexec = nullptr; }
        if (!(key == pl->gkey)) pl->gfailed = false;
        pl->gkey = key;
        return dist_enqueue<T>(pl, alpha, d_data_local, d_x_local, nrhs, beta, d_y_own);
    }
    const i64 l0 = h->launches;
    cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { cudaGetLastError(); pl->gfailed = true; return dist_enqueue<T>(pl, alpha, d_data_local, d_x_local, nrhs, beta, d_y_own); }
    int32_t rc = dist_enqueue<T>(pl, alpha, d_data_local, d_x_local, nrhs, beta, d_y_own);
    cudaGraph_t graph = nullptr;
    e = cudaStreamEndCapture(h->stream, &graph);
    const i64 captured = h->launches - l0;
    h->launches = l0;
    if (rc != BBM_OK || e != cudaSuccess || !graph) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        pl->gfailed = true;
        if (rc != BBM_OK) return rc;
        return dist_enqueue<T>(pl, alpha, d_data_local, d_x_local, nrhs, beta, d_y_own);
    }
    e = cudaGraphInstantiate(&pl->gexec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        cudaGetLastError();
        pl->gexec = nullptr; pl->gfailed = true;
        return dist_enqueue<T>(pl, alpha, d_data_local, d_x_local, nrhs, beta, d_y_own);
    }
    pl->glaunches = captured;
    BBM_CUDA(h, cudaGraphLaunch(pl->gexec, h->stream));
    h->launches += pl->glaunches;
    return BBM_OK;
}


// y_own (host) <- alpha * A[K0:K1,:] * x + beta * y_own with the rank's x_own in (pinned) host memory.
// Everything a rank does for one product is pipelined over block-row chunks on four streams:
//   copy_in : [blocks the neighbours need -> staging] [x_own chunk 0] [x_own chunk 1] ...
//   comm    :        NCCL send (staging) / recv (halo parts of x_local)
//   compute :                         [chunk 0: waits x chunk 0 (+ halo)] [chunk 1] ...   (sub-views of the local slab)
//   copy_out:                                              [y chunk 0] [y chunk 1] ...
// so the call costs about max(upload, download) instead of upload + product + download.
template <typename T>
int32_t dist_mul_host_impl(bbm_dist_bbb_plan* pl, T alpha, const T* d_data_local, const T* h_x_own, T beta, T* h_y_own) {
    if (!plan_valid(pl)) return BBM_ERR_HANDLE;
    bbm_handle_t h = pl->h;
    bbm_comm* cm = pl->comm;
    const bbm_dist_layout& L = pl->lay;
    if ((L.n_own > 0 && !h_x_own) || (L.m_own > 0 && !h_y_own)) return bbm_fail(h, BBM_ERR_ARG, "null host pointer");
    BBM_CUDA(h, cudaSetDevice(h->device));
    const bool exchange = cm && cm->comm && cm->nranks > 1 && (!pl->sends.empty() || !pl->recvs.empty());
    if (pl->ws_elem != (int)sizeof(T)) {
        for (void** w : {&pl->ws_x, &pl->ws_y, &pl->ws_s})
            if (*w) { cudaFree(*w); *w = nullptr; }
        i64 nsend = 0;
        for (size_t i = 0; i + 2 < pl->sends.size(); i += 3) nsend += pl->sends[i + 2];
        cudaError_t e = cudaMalloc(&pl->ws_x, (size_t)std::max<i64>(1, L.n_local) * sizeof(T) + 64);
        if (e == cudaSuccess) e = cudaMalloc(&pl->ws_y, (size_t)std::max<i64>(1, L.m_own) * sizeof(T) + 64);
        if (e == cudaSuccess) e = cudaMalloc(&pl->ws_s, (size_t)std::max<i64>(1, nsend) * sizeof(T) + 64);
        if (e != cudaSuccess) { cudaGetLastError(); return bbm_fail(h, BBM_ERR_ALLOC, "host-vector workspaces: %s", cudaGetErrorString(e)); }
        pl->ws_elem = (int)sizeof(T);
    }
    if (pl->hchunks.empty() && L.K1 > L.K0) {
        const i64 nown = L.K1 - L.K0;
        const int nch = (int)std::max<i64>(1, std::min<i64>(std::min<i64>(bbm_opt(h, "bbb.host_chunks", 8), 32), nown));
        std::vector<i64> kb(nch + 1);
        if (bbm_bbb_partition_rows(nown, pl->r.data() + L.K0, nch, kb.data()) != BBM_OK) return bbm_fail(h, BBM_ERR_ARG, "bad block sizes");
        pl->C_local.assign(L.Jhi - L.Jlo + 1, 0);
        for (i64 J = L.Jlo; J < L.Jhi; ++J) pl->C_local[J - L.Jlo + 1] = pl->C_local[J - L.Jlo] + pl->c[J];
        pl->R_own.assign(nown + 1, 0);
        for (i64 K = L.K0; K < L.K1; ++K) pl->R_own[K - L.K0 + 1] = pl->R_own[K - L.K0] + pl->r[K];
        for (int i = 0; i < nch; ++i) {
            if (kb[i + 1] <= kb[i]) continue;
            bbm_dist_bbb_plan::HostChunk ch{L.K0 + kb[i], L.K0 + kb[i + 1], 0, 0, false, nullptr};
            ch.Jl = clampi(ch.Ka - pl->l, L.Jlo, L.Jhi);
            ch.Jh = clampi(ch.Kb - 1 + pl->u + 1, ch.Jl, L.Jhi);
            if (pl->P == 0) ch.Jh = ch.Jl;
            ch.halo = ch.Jl < L.J0 || ch.Jh > L.J1;
            const i64 sft = ch.Ka - ch.Jl;
            int32_t rc = bbm_bbb_desc_create(h, ch.Kb - ch.Ka, ch.Jh - ch.Jl, pl->r.data() + ch.Ka, pl->c.data() + ch.Jl, pl->l - sft, pl->u + sft,
                                             pl->lam, pl->mu, &ch.sub);
            if (rc != BBM_OK) return rc;
            pl->hchunks.push_back(ch);
        }
        pl->hev.resize(2 * pl->hchunks.size() + 3, nullptr);
        for (auto& e : pl->hev) BBM_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (!h->copy_in) {
        BBM_CUDA(h, cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking));
        BBM_CUDA(h, cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking));
    }
    T* dX = static_cast<T*>(pl->ws_x);
    T* dY = static_cast<T*>(pl->ws_y);
    T* dS = static_cast<T*>(pl->ws_s);
    const size_t nch = pl->hchunks.size();
    cudaEvent_t ev0 = nch ? pl->hev[2 * nch] : nullptr, ev_stage = nch ? pl->hev[2 * nch + 1] : nullptr, ev_halo = nch ? pl->hev[2 * nch + 2] : nullptr;
    if (nch) {
        BBM_CUDA(h, cudaEventRecord(ev0, h->stream));            // earlier work on the handle's stream
        BBM_CUDA(h, cudaStreamWaitEvent(h->copy_in, ev0, 0));
        BBM_CUDA(h, cudaStreamWaitEvent(h->copy_out, ev0, 0));
    }
    const i64 own0 = L.own_offset;
    if (exchange) {
        Nccl& n = nccl();
        // the blocks the neighbours need go up first, into a staging buffer of their own (the owned part of x_local is
        // filled chunk by chunk below; the halo parts only by the receives)
        i64 so = 0;
        for (size_t i = 0; i + 2 < pl->sends.size(); i += 3) {
            const i64 off = pl->sends[i + 1], cnt = pl->sends[i + 2];
            BBM_CUDA(h, cudaMemcpyAsync(dS + so, h_x_own + (off - own0), (size_t)cnt * sizeof(T), cudaMemcpyHostToDevice, h->copy_in ? h->copy_in : h->stream));
            so += cnt;
        }
        cudaStream_t cin = h->copy_in ? h->copy_in : h->stream;
        cudaEvent_t evs = ev_stage ? ev_stage : cm->ev_ready;
        BBM_CUDA(h, cudaEventRecord(evs, cin));
        BBM_CUDA(h, cudaStreamWaitEvent(cm->stream, evs, 0));
        if (ev0) BBM_CUDA(h, cudaStreamWaitEvent(cm->stream, ev0, 0));
        ncclResult_t nrc = n.GroupStart();
        for (size_t i = 0; i + 2 < pl->recvs.size() && nrc == ncclSuccess; i += 3)
            nrc = n.Recv(dX + pl->recvs[i + 1], (size_t)pl->recvs[i + 2] * sizeof(T), ncclChar, (int)pl->recvs[i], cm->comm, cm->stream);
        so = 0;
        for (size_t i = 0; i + 2 < pl->sends.size() && nrc == ncclSuccess; i += 3) {
            nrc = n.Send(dS + so, (size_t)pl->sends[i + 2] * sizeof(T), ncclChar, (int)pl->sends[i], cm->comm, cm->stream);
            so += pl->sends[i + 2];
        }
        ncclResult_t nrc2 = n.GroupEnd();
        if (nrc == ncclSuccess) nrc = nrc2;
        if (nrc != ncclSuccess) return bbm_fail(h, BBM_ERR_NCCL, "halo exchange: %s", n.GetErrorString(nrc));
        if (ev_halo) BBM_CUDA(h, cudaEventRecord(ev_halo, cm->stream));
    }
    const i64 ldx = std::max<i64>(1, L.n_local), ldy = std::max<i64>(1, L.m_own);
    i64 up = own0;  // x_local[own0, up) is on the device
    const i64 own1 = own0 + L.n_own;
    int32_t rc = BBM_OK;
    for (size_t i = 0; i < nch; ++i) {
        const auto& ch = pl->hchunks[i];
        const i64 need = std::min<i64>(own1, std::max<i64>(own0, pl->C_local[ch.Jh - L.Jlo]));
        if (need > up) {
            BBM_CUDA(h, cudaMemcpyAsync(dX + up, h_x_own + (up - own0), (size_t)(need - up) * sizeof(T), cudaMemcpyHostToDevice, h->copy_in));
            up = need;
        }
        const i64 r0 = pl->R_own[ch.Ka - L.K0], nr = pl->R_own[ch.Kb - L.K0] - r0;
        if (beta != T(0) && nr > 0)
            BBM_CUDA(h, cudaMemcpyAsync(dY + r0, h_y_own + r0, (size_t)nr * sizeof(T), cudaMemcpyHostToDevice, h->copy_in));
        BBM_CUDA(h, cudaEventRecord(pl->hev[2 * i], h->copy_in));
        BBM_CUDA(h, cudaStreamWaitEvent(h->stream, pl->hev[2 * i], 0));
        if (ch.halo && exchange) BBM_CUDA(h, cudaStreamWaitEvent(h->stream, ev_halo, 0));
        const i64 c0 = pl->C_local[ch.Jl - L.Jlo];
        if (sizeof(T) == 8)
            rc = bbm_bbb_mul_f64(h, ch.sub, (double)alpha, (const double*)d_data_local + pl->P * c0, (const double*)dX + c0, ldx, 1, (double)beta,
                                 (double*)dY + r0, ldy);
        else
            rc = bbm_bbb_mul_f32(h, ch.sub, (float)alpha, (const float*)d_data_local + pl->P * c0, (const float*)dX + c0, ldx, 1, (float)beta,
                                 (float*)dY + r0, ldy);
        if (rc != BBM_OK) return rc;
        BBM_CUDA(h, cudaEventRecord(pl->hev[2 * i + 1], h->stream));
        BBM_CUDA(h, cudaStreamWaitEvent(h->copy_out, pl->hev[2 * i + 1], 0));
        if (nr > 0) BBM_CUDA(h, cudaMemcpyAsync(h_y_own + r0, dY + r0, (size_t)nr * sizeof(T), cudaMemcpyDeviceToHost, h->copy_out));
    }
    if (nch) {
        if (up < own1)   // owned columns no owned row touches (they still had to reach the neighbours above)
            up = own1;
        BBM_CUDA(h, cudaStreamSynchronize(h->copy_out));
        BBM_CUDA(h, cudaStreamSynchronize(h->copy_in));
    }
    if (exchange) BBM_CUDA(h, cudaStreamSynchronize(cm->stream));
    BBM_CUDA(h, cudaStreamSynchronize(h->stream));
    return BBM_OK;
}

}  // namespace

int32_t bbm_dist_layout_impl(i64 Nr, i64 Nc, const i64* r, const i64* c, i64 l, i64 u, int nranks, int rank, const i64* Kb,
                             bbm_dist_layout* out, std::vector<i64>* sends, std::vector<i64>* recvs) {
    return layout_impl(Nr, Nc, r, c, l, u, nranks, rank, Kb, out, sends, recvs);
}

extern "C" {

int32_t bbm_bbb_partition_rows(int64_t Nr, const int64_t* r, int32_t nranks, int64_t* Kbounds) {
    if (Nr < 0 || nranks < 1 || !Kbounds || (Nr > 0 && !r)) return BBM_ERR_ARG;
    // contiguous block-row ranges balanced by rows (stored entries per block row are proportional
    // to its height); cf. the balanced-remainder split of examples/shared_array_backend.jl:106-107
    std::vector<i64> R(Nr + 1, 0);
    for (i64 K = 0; K < Nr; ++K) { if (r[K] < 0) return BBM_ERR_ARG; R[K + 1] = R[K] + r[K]; }
    const i64 m = R[Nr];
    Kbounds[0] = 0;
    for (int p = 1; p < nranks; ++p) {
        i64 K;
        if (m == 0) {
            K = (Nr * p) / nranks;
        } else {
            const double target = (double)m * p / nranks;
            K = std::lower_bound(R.begin(), R.end(), (i64)(target + 0.5)) - R.begin();
            if (K > 0 && (double)R[K] - target > target - (double)R[K - 1]) --K;  // nearer boundary
        }
        Kbounds[p] = clampi(K, Kbounds[p - 1], Nr);
    }
    Kbounds[nranks] = Nr;
    return BBM_OK;
}

int32_t bbm_bbb_dist_layout(int64_t Nr, int64_t Nc, const int64_t* r, const int64_t* c, int64_t l, int64_t u,
                            int32_t nranks, int32_t rank, const int64_t* Kbounds, bbm_dist_layout* out,
                            int64_t* sendlist, int64_t* recvlist) {
    std::vector<i64> s, v;
    int32_t rc = layout_impl(Nr, Nc, r, c, l, u, nranks, rank, Kbounds, out, &s, &v);
    if (rc != BBM_OK) return rc;
    if (sendlist) std::copy(s.begin(), s.end(), sendlist);
    if (recvlist) std::copy(v.begin(), v.end(), recvlist);
    return BBM_OK;
}

int32_t bbm_comm_unique_id(uint8_t* id128) {
    if (!id128) return BBM_ERR_ARG;
    if (!nccl_load()) return BBM_ERR_NCCL;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != ncclSuccess) return BBM_ERR_NCCL;
    memcpy(id128, id.internal, 128);
    return BBM_OK;
}

int32_t bbm_comm_create(bbm_handle_t h, const uint8_t* id128, int32_t nranks, int32_t rank, bbm_comm_t* out) {
    BBM_REQUIRE_HANDLE(h);
    if (!out || nranks < 1 || rank < 0 || rank >= nranks) return bbm_fail(h, BBM_ERR_ARG, "bad communicator arguments");
    *out = nullptr;
    bbm_comm* c = new (std::nothrow) bbm_comm();
    if (!c) return bbm_fail(h, BBM_ERR_ALLOC, "out of host memory");
    c->h = h; c->nranks = nranks; c->rank = rank;
    BBM_CUDA(h, cudaSetDevice(h->device));
    // id128 == NULL with nranks > 1: a communicator without NCCL -- the caller moves the halo itself
    // (its own MPI / Distributed exchange) and bbm_dist_bbb_mul_* only multiplies.
    if (nranks > 1 && id128) {
        if (!nccl_load()) { delete c; return bbm_fail(h, BBM_ERR_NCCL, "%s", nccl().err.c_str()); }
        ncclUniqueId id;
        memcpy(id.internal, id128, 128);
        ncclResult_t rc = nccl().CommInitRank(&c->comm, nranks, id, rank);
        if (rc != ncclSuccess) { delete c; return bbm_fail(h, BBM_ERR_NCCL, "ncclCommInitRank: %s", nccl().GetErrorString(rc)); }
    }
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    cudaError_t e = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming);
    if (e != cudaSuccess) { bbm_comm_destroy(c); return bbm_fail(h, BBM_ERR_CUDA, "communicator streams: %s", cudaGetErrorString(e)); }
    *out = c;
    return BBM_OK;
}

int32_t bbm_comm_destroy(bbm_comm_t c) {
    if (!c || c->magic != 0xC0221234u) return BBM_ERR_HANDLE;
    if (bbm_valid(c->h)) cudaSetDevice(c->h->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm && nccl().CommDestroy) nccl().CommDestroy(c->comm);
    if (c->ev_ready) cudaEventDestroy(c->ev_ready);
    if (c->ev_halo) cudaEventDestroy(c->ev_halo);
    if (c->stream) cudaStreamDestroy(c->stream);
    c->magic = 0;
    delete c;
    return BBM_OK;
}

int32_t bbm_dist_bbb_plan_create(bbm_comm_t cm, int64_t Nr, int64_t Nc, const int64_t* r, const int64_t* c, int64_t l,
                                 int64_t u, int64_t lam, int64_t mu, const int64_t* Kbounds, bbm_dist_bbb_plan_t* out) {
    if (!comm_valid(cm)) return BBM_ERR_HANDLE;
    bbm_handle_t h = cm->h;
    if (!out) return bbm_fail(h, BBM_ERR_ARG, "null out pointer");
    *out = nullptr;
    std::vector<i64> kb(cm->nranks + 1);
    if (Kbounds) std::copy(Kbounds, Kbounds + cm->nranks + 1, kb.begin());
    else {
        int32_t rc = bbm_bbb_partition_rows(Nr, r, cm->nranks, kb.data());
        if (rc != BBM_OK) return bbm_fail(h, rc, "bad block sizes");
    }
    bbm_dist_bbb_plan* pl = new (std::nothrow) bbm_dist_bbb_plan();
    if (!pl) return bbm_fail(h, BBM_ERR_ALLOC, "out of host memory");
    pl->comm = cm; pl->h = h;
    pl->Nr = Nr; pl->Nc = Nc; pl->l = l; pl->u = u; pl->lam = lam; pl->mu = mu;
    pl->r.assign(r, r + Nr); pl->c.assign(c, c + Nc); pl->kb = kb;
    int32_t rc = layout_impl(Nr, Nc, r, c, l, u, cm->nranks, cm->rank, kb.data(), &pl->lay, &pl->sends, &pl->recvs);
    if (rc != BBM_OK) { delete pl; return bbm_fail(h, rc, "bad partition"); }
    const i64 nb = std::max<i64>(0, l + u + 1), w = std::max<i64>(0, lam + mu + 1);
    pl->P = nb * w;
    const bbm_dist_layout& L = pl->lay;
    std::vector<i64> C(Nc + 1, 0), R(Nr + 1, 0);
    for (i64 J = 0; J < Nc; ++J) C[J + 1] = C[J] + c[J];
    for (i64 K = 0; K < Nr; ++K) R[K + 1] = R[K] + r[K];
    // pieces: interior [Ka,Kb), left boundary [K0,Ka), right boundary [Kb,K1)
    const i64 lo[3] = {L.Ka, L.K0, L.Kb}, hi[3] = {L.Kb, L.Ka, L.K1};
    for (int i = 0; i < 3; ++i) {
        if (hi[i] <= lo[i]) continue;
        // block columns of the piece, kept inside the local column range
        i64 Jl = clampi(lo[i] - l, L.Jlo, L.Jhi), Jh = clampi(hi[i] + u, Jl, L.Jhi);
        if (pl->P == 0) { Jl = L.Jlo; Jh = L.Jlo; }
        const i64 s = lo[i] - Jl;  // sub-view shift: (l,u) -> (l-s, u+s)
        rc = bbm_bbb_desc_create(h, hi[i] - lo[i], Jh - Jl, r + lo[i], c + Jl, l - s, u + s, lam, mu, &pl->desc[i]);
        if (rc != BBM_OK) { bbm_dist_bbb_plan_destroy(pl); return rc; }
        pl->coloff[i] = C[Jl] - C[L.Jlo];
        pl->rowoff[i] = R[lo[i]] - R[L.K0];
    }
    *out = pl;
    return BBM_OK;
}

int32_t bbm_dist_bbb_plan_destroy(bbm_dist_bbb_plan_t pl) {
    if (!pl || pl->magic != 0xD157B200u) return BBM_ERR_HANDLE;
    for (int i = 0; i < 3; ++i)
        if (pl->desc[i]) bbm_bbb_desc_destroy(pl->desc[i]);
    if (pl->gexec) cudaGraphExecDestroy(pl->gexec);
    if (pl->desc_all) bbm_bbb_desc_destroy(pl->desc_all);
    for (int i = 0; i < 2; ++i)
        if (pl->ipc_sig[i]) cudaIpcCloseMemHandle(pl->ipc_sig[i]);
    if (pl->sig) cudaFree(pl->sig);
    if (pl->err_host) cudaFreeHost(pl->err_host);
    for (auto& ch : pl->hchunks)
        if (ch.sub) bbm_bbb_desc_destroy(ch.sub);
    for (auto& e : pl->hev)
        if (e) cudaEventDestroy(e);
    if (pl->ws_x) cudaFree(pl->ws_x);
    if (pl->ws_y) cudaFree(pl->ws_y);
    if (pl->ws_s) cudaFree(pl->ws_s);
    pl->magic = 0;
    delete pl;
    return BBM_OK;
}

int32_t bbm_dist_bbb_plan_layout(bbm_dist_bbb_plan_t pl, bbm_dist_layout* out) {
    if (!plan_valid(pl) || !out) return BBM_ERR_HANDLE;
    *out = pl->lay;
    return BBM_OK;
}


int32_t bbm_dist_bbb_enable_peer(bbm_dist_bbb_plan_t pl, int32_t elem_size) {
    if (!plan_valid(pl)) return BBM_ERR_HANDLE;
    bbm_handle_t h = pl->h;
    bbm_comm* cm = pl->comm;
    if (!cm || !cm->comm || cm->nranks < 2) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "fused halo push needs an NCCL communicator over >= 2 ranks");
    if (elem_size != 4 && elem_size != 8) return bbm_fail(h, BBM_ERR_ARG, "element size must be 4 or 8");
    BBM_CUDA(h, cudaSetDevice(h->device));
    const int rank = cm->rank, nranks = cm->nranks;
    const bbm_dist_layout& L = pl->lay;
    // halo element counts (left, right) of any rank, from the deterministic layout
    auto halo_counts = [&](int q, i64* cntL, i64* cntR, i64* colL, i64* colR) -> bool {
        bbm_dist_layout Lq;
        std::vector<i64> sq, rq;
        if (layout_impl(pl->Nr, pl->Nc, pl->r.data(), pl->c.data(), pl->l, pl->u, nranks, q, pl->kb.data(), &Lq, &sq, &rq) != BBM_OK) return false;
        *cntL = *cntR = 0; *colL = *colR = 0;
        bool fine = true;
        for (size_t k = 0; k + 2 < rq.size(); k += 3) {
            const int from = (int)rq[k];
            if (from == q - 1 && *cntL == 0 && rq[k + 1] < Lq.own_offset) { *cntL = rq[k + 2]; *colL = rq[k + 1]; }
            else if (from == q + 1 && *cntR == 0 && rq[k + 1] >= Lq.own_offset + Lq.n_own) { *cntR = rq[k + 2]; *colR = rq[k + 1]; }
            else fine = false;  // a halo that reaches past the immediate neighbour
        }
        return fine;
    };
    auto pad256 = [](size_t b) { return (b + 255) & ~size_t(255); };
    // device block of rank q: [sig 256 B][pad 256][L0][L1][R0][R1], each buffer padded by 256 B on both sides
    auto buf_off = [&](i64 cntL, i64 cntR, int side, int parity) -> size_t {
        const size_t bl = pad256((size_t)cntL * elem_size) + 256, br = pad256((size_t)cntR * elem_size) + 256;
        size_t off = 512;
        if (side == 0) return off + (size_t)parity * bl;
        off += 2 * bl;
        return off + (size_t)parity * br;
    };
    i64 myL = 0, myR = 0, mycolL = 0, mycolR = 0;
    bool ok = pl->P > 0 && pl->l + pl->u + 1 <= 8 && L.m_own > 0;
    ok = halo_counts(rank, &myL, &myR, &mycolL, &mycolR) && ok;
    BbmPeerParams tpl;
    tpl.own_J0 = (int)(L.J0 - L.Jlo);
    tpl.own_J1 = (int)(L.J1 - L.Jlo);
    tpl.has_halo[0] = myL > 0; tpl.has_halo[1] = myR > 0;
    tpl.halo_col0[0] = mycolL; tpl.halo_col0[1] = mycolR;
    i64 nbL[2] = {0, 0}, nbR[2] = {0, 0};  // halo counts of my left / right neighbour
    for (size_t i = 0; i + 2 < pl->sends.size(); i += 3) {
        const int q = (int)pl->sends[i];
        const int sd = q == rank - 1 ? 0 : (q == rank + 1 ? 1 : -1);
        if (sd < 0 || tpl.cnt[sd] != 0) { ok = false; continue; }
        tpl.src_off[sd] = pl->sends[i + 1];
        tpl.cnt[sd] = pl->sends[i + 2];
        i64 cl, cr;
        if (!halo_counts(q, &nbL[sd], &nbR[sd], &cl, &cr)) ok = false;
        // what I send to my left neighbour is its RIGHT halo and vice versa
        if ((sd == 0 ? nbR[sd] : nbL[sd]) != tpl.cnt[sd]) ok = false;
    }
    for (int sd = 0; sd < 2; ++sd) {  // neighbours I only receive from still need their signal block
        const int q = sd == 0 ? rank - 1 : rank + 1;
        if (tpl.has_halo[sd] && tpl.cnt[sd] == 0 && q >= 0 && q < nranks) {
            i64 cl, cr;
            if (!halo_counts(q, &nbL[sd], &nbR[sd], &cl, &cr)) ok = false;
        }
    }
    // one descriptor for all owned rows over the local column slab
    if (ok && !pl->desc_all) {
        const i64 s = L.K0 - L.Jlo;
        int32_t rc = bbm_bbb_desc_create(h, L.K1 - L.K0, L.Jhi - L.Jlo, pl->r.data() + L.K0, pl->c.data() + L.Jlo, pl->l - s,
                                         pl->u + s, pl->lam, pl->mu, &pl->desc_all);
        if (rc != BBM_OK) ok = false;
    }
    // From here to the end every rank takes part in two all-gathers: a local failure must not return early (the
    // peers would hang in the collective) -- it is folded into the ok flag that is gathered.
    const char* local_fail = nullptr;
    auto soft = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess) { cudaGetLastError(); ok = false; if (!local_fail) local_fail = what; }
    };
    pl->peer_ok = false;
    // the old block may still be mapped by the neighbours: they close their mappings below before anybody pushes
    // again, and nobody pushes while peer_ok is false; the block itself is released only after the first all-gather
    BbmPeerSig* old_sig = pl->sig;
    pl->sig = nullptr;
    const size_t block_bytes = buf_off(myL, myR, 1, 1) + pad256((size_t)myR * elem_size) + 512;
    soft(cudaMalloc(reinterpret_cast<void**>(&pl->sig), block_bytes), "cudaMalloc(halo block)");
    if (pl->sig) soft(cudaMemset(pl->sig, 0, block_bytes), "cudaMemset(halo block)");
    if (!pl->err_host) {
        soft(cudaHostAlloc(reinterpret_cast<void**>(&pl->err_host), 64, cudaHostAllocMapped), "cudaHostAlloc(error word)");
        if (pl->err_host) soft(cudaHostGetDevicePointer(reinterpret_cast<void**>(&pl->err_host_dev), pl->err_host, 0), "cudaHostGetDevicePointer");
    }
    if (pl->err_host) *pl->err_host = 0;
    soft(cudaDeviceGetAttribute(&pl->clock_khz, cudaDevAttrClockRate, h->device), "cudaDeviceGetAttribute(clock rate)");
    soft(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    // all-gather (ok flag, IPC handle of the block) over NCCL
    struct Rec { int ok; int pad[3]; cudaIpcMemHandle_t blk; };
    Rec mine{};
    if (!pl->sig || cudaIpcGetMemHandle(&mine.blk, pl->sig) != cudaSuccess) { cudaGetLastError(); ok = false; }
    std::vector<Rec> all(nranks);
    Rec* d_all = nullptr;
    // the gather buffer itself: without it this rank cannot take part at all (the one unavoidable early return)
    if (cudaMalloc(&d_all, sizeof(Rec) * nranks) != cudaSuccess) { cudaGetLastError(); return bbm_fail(h, BBM_ERR_ALLOC, "cudaMalloc(all-gather buffer)"); }
    mine.ok = ok ? 1 : 0;
    soft(cudaMemcpy(d_all + rank, &mine, sizeof(Rec), cudaMemcpyHostToDevice), "cudaMemcpy(record)");
    ncclResult_t nrc = nccl().AllGather(d_all + rank, d_all, sizeof(Rec), ncclChar, cm->comm, cm->stream);
    if (nrc != ncclSuccess) { cudaFree(d_all); if (old_sig) cudaFree(old_sig); return bbm_fail(h, BBM_ERR_NCCL, "ncclAllGather: %s", nccl().GetErrorString(nrc)); }
    soft(cudaStreamSynchronize(cm->stream), "cudaStreamSynchronize");
    soft(cudaMemcpy(all.data(), d_all, sizeof(Rec) * nranks, cudaMemcpyDeviceToHost), "cudaMemcpy(records)");
    // every rank has passed the gather: nobody reads the previous blocks any more
    for (int sd = 0; sd < 2; ++sd)
        if (pl->ipc_sig[sd]) { cudaIpcCloseMemHandle(pl->ipc_sig[sd]); pl->ipc_sig[sd] = nullptr; }
    bool all_ok = ok;
    for (int q = 0; q < nranks; ++q) all_ok = all_ok && all[q].ok == 1;
    bool opened = all_ok;
    for (int sd = 0; sd < 2; ++sd) {
        const int q = sd == 0 ? rank - 1 : rank + 1;
        if (!all_ok || q < 0 || q >= nranks || (tpl.cnt[sd] == 0 && !tpl.has_halo[sd])) continue;
        cudaError_t e = cudaIpcOpenMemHandle(&pl->ipc_sig[sd], all[q].blk, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {  // no peer mapping on this system: every rank must learn it (second all-gather below)
            cudaGetLastError();
            bbm_fail(h, BBM_ERR_UNSUPPORTED, "cudaIpcOpenMemHandle(rank %d): %s", q, cudaGetErrorString(e));
            pl->ipc_sig[sd] = nullptr;
            opened = false;
            continue;
        }
        tpl.remote[sd] = static_cast<BbmPeerSig*>(pl->ipc_sig[sd]);
    }
    for (int sd = 0; sd < 2; ++sd) {
        tpl.ack_wait[sd] = tpl.cnt[sd] > 0 && !tpl.has_halo[sd];
        tpl.ack_send[sd] = tpl.has_halo[sd] && tpl.cnt[sd] == 0;
    }
    tpl.local = pl->sig;
    pl->peer_tpl = tpl;
    pl->reg_elem = elem_size;
    pl->epoch = 0;
    // byte offsets of the buffers this rank reads (its own) and writes (the neighbours')
    for (int par = 0; par < 2; ++par) {
        pl->own_buf[0][par] = buf_off(myL, myR, 0, par) + 256;
        pl->own_buf[1][par] = buf_off(myL, myR, 1, par) + 256;
        pl->nb_buf[0][par] = buf_off(nbL[0], nbR[0], 1, par) + 256;  // left neighbour's RIGHT halo
        pl->nb_buf[1][par] = buf_off(nbL[1], nbR[1], 0, par) + 256;  // right neighbour's LEFT halo
    }
    // nobody may push before every rank has reset its flags and opened its mappings; the same all-gather
    // tells every rank whether all mappings could be opened, so that the ranks switch paths together
    mine.ok = opened ? 1 : 0;
    if (cudaMemcpy(d_all + rank, &mine, sizeof(Rec), cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); cudaMemset(d_all + rank, 0, sizeof(Rec)); opened = false; }
    nrc = nccl().AllGather(d_all + rank, d_all, sizeof(Rec), ncclChar, cm->comm, cm->stream);
    cudaStreamSynchronize(cm->stream);
    if (nrc == ncclSuccess && cudaMemcpy(all.data(), d_all, sizeof(Rec) * nranks, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); opened = false; }
    cudaFree(d_all);
    if (old_sig) cudaFree(old_sig);   // every neighbour closed its mapping of it before this second gather
    if (nrc != ncclSuccess) return bbm_fail(h, BBM_ERR_NCCL, "enable_peer barrier: %s", nccl().GetErrorString(nrc));
    for (int q = 0; q < nranks; ++q) opened = opened && all[q].ok == 1;
    if (!opened) {
        for (int sd = 0; sd < 2; ++sd)
            if (pl->ipc_sig[sd]) { cudaIpcCloseMemHandle(pl->ipc_sig[sd]); pl->ipc_sig[sd] = nullptr; }
        if (local_fail) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "fused halo push unavailable (%s failed on this rank); the NCCL exchange stays in use", local_fail);
        if (!all_ok) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "fused halo push not applicable to this partition; the NCCL exchange stays in use");
        return bbm_fail(h, BBM_ERR_UNSUPPORTED, "peer memory mappings are not available on every rank; the NCCL exchange stays in use");
    }
    pl->peer_ok = true;
    return BBM_OK;
}

int32_t bbm_dist_bbb_peer_status(bbm_dist_bbb_plan_t pl, int32_t* active, int32_t* timed_out) {
    if (!plan_valid(pl)) return BBM_ERR_HANDLE;
    bbm_handle_t h = pl->h;
    if (active) *active = pl->peer_ok ? 1 : 0;
    if (timed_out) {
        *timed_out = 0;
        if (pl->sig) {
            BBM_CUDA(h, cudaSetDevice(h->device));
            BBM_CUDA(h, cudaStreamSynchronize(h->stream));
            BbmPeerSig sg;
            BBM_CUDA(h, cudaMemcpy(&sg, pl->sig, sizeof sg, cudaMemcpyDeviceToHost));
            *timed_out = (int32_t)(sg.err != 0 || (pl->err_host && *reinterpret_cast<volatile unsigned*>(pl->err_host) != 0));
        }
    }
    return BBM_OK;
}

int32_t bbm_dist_bbb_mul_f64(bbm_dist_bbb_plan_t pl, double alpha, const double* d_data_local, double* d_x_local,
                             int64_t nrhs, double beta, double* d_y_own) {
    return dist_mul_impl<double>(pl, alpha, d_data_local, d_x_local, nrhs, beta, d_y_own);
}
int32_t bbm_dist_bbb_mul_f32(bbm_dist_bbb_plan_t pl, float alpha, const float* d_data_local, float* d_x_local,
                             int64_t nrhs, float beta, float* d_y_own) {
    return dist_mul_impl<float>(pl, alpha, d_data_local, d_x_local, nrhs, beta, d_y_own);
}
int32_t bbm_dist_bbb_mul_host_f64(bbm_dist_bbb_plan_t pl, double alpha, const double* d_data_local, const double* h_x_own, double beta,
                                  double* h_y_own) { return dist_mul_host_impl<double>(pl, alpha, d_data_local, h_x_own, beta, h_y_own); }
int32_t bbm_dist_bbb_mul_host_f32(bbm_dist_bbb_plan_t pl, float alpha, const float* d_data_local, const float* h_x_own, float beta,
                                  float* h_y_own) { return dist_mul_host_impl<float>(pl, alpha, d_data_local, h_x_own, beta, h_y_own); }

}  // extern "C"
