// Row-block-partitioned BlockBandedMatrix over several B200s (one process per GPU): matvec and A*B (cfg4).
//
// Reference precedent: the row-block data parallelism of examples/shared_array_backend.jl:92-122.  A contiguous
// block-row range [S0,S1) of a BlockBandedMatrix is a BlockSkylineMatrix with per-column bandwidths (block column J
// keeps the block rows colsupport(J) & [S0,S1); cf. the sub-view bandwidths of src/linalg.jl:78-87).  Unlike the
// banded-block-banded format its storage is NOT a slice of the parent (every panel is cut by rows), so each rank holds
// its own panels; the local products are the single-GPU kernels on those local ragged structures.
//
//   matvec   y[K0:K1) = A[K0:K1,:] x      x is owned like y; the x block columns inside the l+u halo travel by NCCL
//                                          send/recv on the communicator's stream while the interior block rows
//                                          (which touch owned columns only) are already being multiplied.
//   product  C[K0:K1,:] = A[K0:K1,:] B    B's block rows are owned like x; rank p stores the rows [S0,S1) its A rows
//                                          multiply with.  The halo rows are packed block by block into one buffer per
//                                          neighbour (gather kernel), sent with NCCL, and scattered into the local
//                                          panels; the tiles of C that only multiply owned rows of B run meanwhile.
// The planning functions are pure host code (no GPU, no handle): the CPU tests drive them.
#include <algorithm>
#include <new>

#include "bbm_internal.h"
#include "dist_internal.h"

namespace {

inline i64 clampi(i64 v, i64 lo, i64 hi) { return v < lo ? lo : (v > hi ? hi : v); }

// local structure of block rows [S0,S1) of a (Nr x Nc)-block matrix with constant bandwidths (l,u)
void rowrange_structure(i64 Nr, i64 Nc, i64 l, i64 u, i64 S0, i64 S1, i64* Jlo, i64* Jhi, std::vector<i64>* lloc, std::vector<i64>* uloc) {
    lloc->clear(); uloc->clear();
    if (S1 <= S0 || l + u + 1 <= 0) {
        *Jlo = *Jhi = clampi(S0, 0, Nc);
        return;
    }
    *Jlo = clampi(S0 - l, 0, Nc);
    *Jhi = clampi(S1 - 1 + u + 1, *Jlo, Nc);
    for (i64 J = *Jlo; J < *Jhi; ++J) {
        const i64 lo = std::max<i64>(std::max<i64>(J - u, 0), S0) - S0;
        const i64 hi = std::min<i64>(std::min<i64>(J + l, Nr - 1), S1 - 1) - S0;
        const i64 Jp = J - *Jlo;
        if (hi < lo) { lloc->push_back(-1 - Jp); uloc->push_back(Jp); }   // empty column
        else { lloc->push_back(hi - Jp); uloc->push_back(Jp - lo); }
    }
}

void owned_cols(const i64* Kb, int nranks, i64 Nc, int p, i64* J0, i64* J1) {
    *J0 = std::min(Kb[p], Nc);
    *J1 = (p == nranks - 1) ? Nc : std::min(Kb[p + 1], Nc);
    if (*J1 < *J0) *J1 = *J0;
}

struct BlockRef { i64 N, J; };

struct MmLayout {
    i64 K0, K1, S0, S1, O0, O1;
    std::vector<std::vector<BlockRef>> sends, recvs;   // per peer, in ascending (N, J) order on both sides
};

int32_t mm_layout(i64 NrA, i64 NcA, i64 NcB, const i64* cA, const i64* cB, i64 lA, i64 uA, i64 lB, i64 uB, int nranks, int rank,
                  const i64* Kb, MmLayout* out) {
    if (NrA < 0 || NcA < 0 || NcB < 0 || nranks < 1 || rank < 0 || rank >= nranks || !Kb || !out) return BBM_ERR_ARG;
    if (Kb[0] != 0 || Kb[nranks] != NrA) return BBM_ERR_ARG;
    for (int p = 0; p < nranks; ++p)
        if (Kb[p + 1] < Kb[p]) return BBM_ERR_ARG;
    auto stored = [&](int p, i64* s0, i64* s1) {   // B block rows rank p multiplies with = column hull of its A rows
        std::vector<i64> a, b;
        rowrange_structure(NrA, NcA, lA, uA, Kb[p], Kb[p + 1], s0, s1, &a, &b);
    };
    out->K0 = Kb[rank]; out->K1 = Kb[rank + 1];
    stored(rank, &out->S0, &out->S1);
    owned_cols(Kb, nranks, NcA, rank, &out->O0, &out->O1);
    out->sends.assign(nranks, {}); out->recvs.assign(nranks, {});
    for (int q = 0; q < nranks; ++q) {
        if (q == rank) continue;
        i64 q0, q1, t0, t1;
        owned_cols(Kb, nranks, NcA, q, &q0, &q1);
        stored(q, &t0, &t1);
        for (int dir = 0; dir < 2; ++dir) {
            // dir 0: what I need and q owns; dir 1: what q needs and I own
            const i64 lo = dir == 0 ? std::max(out->S0, q0) : std::max(t0, out->O0);
            const i64 hi = dir == 0 ? std::min(out->S1, q1) : std::min(t1, out->O1);
            auto& list = dir == 0 ? out->recvs[q] : out->sends[q];
            for (i64 N = lo; N < hi; ++N) {
                if (cA[N] == 0) continue;
                for (i64 J = std::max<i64>(N - lB, 0); J <= std::min<i64>(N + uB, NcB - 1); ++J)
                    if (cB[J] > 0) list.push_back(BlockRef{N, J});
            }
        }
    }
    return BBM_OK;
}

// ---- pack / unpack of halo blocks ------------------------------------------------------------------------------
struct PackEnt { i64 start, ld, boff; int32_t rows, cols; };   // block in the skyline storage <-> contiguous rows x cols in the buffer

template <typename T, bool PACK>
__global__ void __launch_bounds__(256) sky_pack_kernel(T* __restrict__ data, T* __restrict__ buf, const PackEnt* __restrict__ ents) {
    const PackEnt e = ents[blockIdx.x];
    const i64 n = (i64)e.rows * e.cols;
    for (i64 idx = (i64)blockIdx.y * 256 + threadIdx.x; idx < n; idx += (i64)gridDim.y * 256) {
        const i64 j = idx / e.rows, i = idx - j * e.rows;
        if (PACK) buf[e.boff + idx] = data[e.start + i + j * e.ld];
        else data[e.start + i + j * e.ld] = buf[e.boff + idx];
    }
}

}  // namespace

struct bbm_dist_sky_plan {
    uint32_t magic = 0xD157517Au;
    bbm_comm* comm = nullptr;
    bbm_handle_t h = nullptr;
    bbm_dist_layout lay{};
    std::vector<i64> sends, recvs;       // (peer, x_local offset, count) triples
    i64 Jlo = 0, Jhi = 0;                // block columns of the local skyline structure
    std::vector<i64> lloc, uloc;
    i64 x_off = 0;                       // where the local structure's columns start in x_local
    i64 n_cols = 0, numel = 0;
    bbm_sky_desc_t desc = nullptr;
};

struct bbm_dist_sky_mm_plan {
    uint32_t magic = 0xD1575A33u;
    bbm_comm* comm = nullptr;
    bbm_handle_t h = nullptr;
    MmLayout lay;
    i64 Jlo[3] = {0, 0, 0}, Jhi[3] = {0, 0, 0}, numel[3] = {0, 0, 0};   // A_loc, B_loc, C_loc
    std::vector<i64> lloc[3], uloc[3];
    bbm_sky_desc_t desc[3] = {nullptr, nullptr, nullptr};
    // per peer: pack tables and element counts (send / recv)
    struct Peer { int q = 0; PackEnt* d_ents = nullptr; int nents = 0; i64 count = 0; i64 maxblk = 0; void* buf = nullptr; };
    std::vector<Peer> psend, precv;
    int buf_elem = 0;
};

namespace {

inline bool plan_valid(bbm_dist_sky_plan* p) { return p && p->magic == 0xD157517Au && bbm_valid(p->h); }
inline bool mm_valid(bbm_dist_sky_mm_plan* p) { return p && p->magic == 0xD1575A33u && bbm_valid(p->h); }

template <typename T>
int32_t dist_sky_mul_impl(bbm_dist_sky_plan* pl, T alpha, const T* d_data_local, T* d_x_local, i64 nrhs, T beta, T* d_y_own) {
    if (!plan_valid(pl)) return BBM_ERR_HANDLE;
    bbm_handle_t h = pl->h;
    bbm_comm* cm = pl->comm;
    if (nrhs < 0) return bbm_fail(h, BBM_ERR_ARG, "nrhs < 0");
    if (nrhs == 0) return BBM_OK;   // (a rank without rows still serves its x blocks to the neighbours)
    BBM_CUDA(h, cudaSetDevice(h->device));
    const bbm_dist_layout& L = pl->lay;
    const i64 ldx = std::max<i64>(1, L.n_local), ldy = std::max<i64>(1, L.m_own);
    const bool exchange = cm && cm->comm && cm->nranks > 1 && (!pl->sends.empty() || !pl->recvs.empty());
    const void* x = d_x_local + pl->x_off;
    const i64 nK = L.K1 - L.K0;
    if (!exchange)
        return bbm_sky_mul_rows(h, pl->desc, (int)sizeof(T), (double)alpha, d_data_local, x, ldx, nrhs, (double)beta, d_y_own, ldy, 0, nK, 0);
    BbmNccl& n = bbm_nccl();
    // the exchange may start once everything queued on the compute stream so far (the producer of x, the previous
    // product's readers of the halo) has finished
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
    BBM_CUDA(h, cudaEventRecord(cm->ev_halo, cm->stream));
    // interior block rows need no halo: multiplied while the exchange is in flight; the boundary rows follow it
    const i64 a = L.Ka - L.K0, b = L.Kb - L.K0;
    int32_t rc = bbm_sky_mul_rows(h, pl->desc, (int)sizeof(T), (double)alpha, d_data_local, x, ldx, nrhs, (double)beta, d_y_own, ldy, a, b, 1);
    if (rc != BBM_OK) return rc;
    BBM_CUDA(h, cudaStreamWaitEvent(h->stream, cm->ev_halo, 0));
    if (a > 0) rc = bbm_sky_mul_rows(h, pl->desc, (int)sizeof(T), (double)alpha, d_data_local, x, ldx, nrhs, (double)beta, d_y_own, ldy, 0, a, 2);
    if (rc == BBM_OK && b < nK) rc = bbm_sky_mul_rows(h, pl->desc, (int)sizeof(T), (double)alpha, d_data_local, x, ldx, nrhs, (double)beta, d_y_own, ldy, b, nK, 2);
    return rc;
}

template <typename T>
int32_t dist_sky_matmul_impl(bbm_dist_sky_mm_plan* pl, T alpha, const T* dA, T* dB, T beta, T* dC) {
    if (!mm_valid(pl)) return BBM_ERR_HANDLE;
    bbm_handle_t h = pl->h;
    bbm_comm* cm = pl->comm;
    BBM_CUDA(h, cudaSetDevice(h->device));
    const MmLayout& L = pl->lay;
    const bool exchange = cm && cm->comm && cm->nranks > 1 && (!pl->psend.empty() || !pl->precv.empty());
    // owned block rows of B in local numbering
    const i64 own_lo = L.O0 - L.S0, own_hi = L.O1 - L.S0;
    if (!exchange)
        return bbm_sky_matmul_phase(h, pl->desc[0], pl->desc[1], pl->desc[2], (int)sizeof(T), (double)alpha, dA, dB, (double)beta, dC, 0, own_lo, own_hi);
    if (pl->buf_elem != (int)sizeof(T)) {   // exchange buffers, sized on first use for this element type
        for (auto* v : {&pl->psend, &pl->precv})
            for (auto& pr : *v) {
                if (pr.buf) { cudaFree(pr.buf); pr.buf = nullptr; }
                cudaError_t e = cudaMalloc(&pr.buf, std::max<size_t>(16, (size_t)pr.count * sizeof(T)));
                if (e != cudaSuccess) { cudaGetLastError(); return bbm_fail(h, BBM_ERR_ALLOC, "halo buffer cudaMalloc: %s", cudaGetErrorString(e)); }
            }
        pl->buf_elem = (int)sizeof(T);
    }
    BbmNccl& n = bbm_nccl();
    BBM_CUDA(h, cudaEventRecord(cm->ev_ready, h->stream));
    BBM_CUDA(h, cudaStreamWaitEvent(cm->stream, cm->ev_ready, 0));
    for (auto& pr : pl->psend) {
        if (pr.nents == 0) continue;
        const unsigned gy = (unsigned)std::min<i64>(64, std::max<i64>(1, (pr.maxblk + 255) / 256));
        sky_pack_kernel<T, true><<<dim3((unsigned)pr.nents, gy), 256, 0, cm->stream>>>(dB, static_cast<T*>(pr.buf), pr.d_ents);
        h->launches += 1;
    }
    ncclResult_t nrc = n.GroupStart();
    for (auto& pr : pl->precv)
        if (nrc == ncclSuccess && pr.count > 0) nrc = n.Recv(pr.buf, (size_t)pr.count * sizeof(T), ncclChar, pr.q, cm->comm, cm->stream);
    for (auto& pr : pl->psend)
        if (nrc == ncclSuccess && pr.count > 0) nrc = n.Send(pr.buf, (size_t)pr.count * sizeof(T), ncclChar, pr.q, cm->comm, cm->stream);
    ncclResult_t nrc2 = n.GroupEnd();
    if (nrc == ncclSuccess) nrc = nrc2;
    if (nrc != ncclSuccess) return bbm_fail(h, BBM_ERR_NCCL, "halo exchange: %s", n.GetErrorString(nrc));
    for (auto& pr : pl->precv) {
        if (pr.nents == 0) continue;
        const unsigned gy = (unsigned)std::min<i64>(64, std::max<i64>(1, (pr.maxblk + 255) / 256));
        sky_pack_kernel<T, false><<<dim3((unsigned)pr.nents, gy), 256, 0, cm->stream>>>(dB, static_cast<T*>(pr.buf), pr.d_ents);
        h->launches += 1;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "halo pack/unpack launch: %s", cudaGetErrorString(e));
    BBM_CUDA(h, cudaEventRecord(cm->ev_halo, cm->stream));
    // tiles of C that only multiply owned rows of B run while the halo rows travel; the others behind their arrival
    int32_t rc = bbm_sky_matmul_phase(h, pl->desc[0], pl->desc[1], pl->desc[2], (int)sizeof(T), (double)alpha, dA, dB, (double)beta, dC, 1, own_lo, own_hi);
    if (rc != BBM_OK) return rc;
    BBM_CUDA(h, cudaStreamWaitEvent(h->stream, cm->ev_halo, 0));
    return bbm_sky_matmul_phase(h, pl->desc[0], pl->desc[1], pl->desc[2], (int)sizeof(T), (double)alpha, dA, dB, (double)beta, dC, 2, own_lo, own_hi);
}

}  // namespace

extern "C" {

int32_t bbm_sky_rowrange_structure(int64_t Nr, int64_t Nc, int64_t l, int64_t u, int64_t S0, int64_t S1, int64_t* Jlo, int64_t* Jhi,
                                   int64_t* lvec, int64_t* uvec) {
    if (Nr < 0 || Nc < 0 || S0 < 0 || S1 > Nr || !Jlo || !Jhi) return BBM_ERR_ARG;
    std::vector<i64> ll, uu;
    rowrange_structure(Nr, Nc, l, u, S0, S1, Jlo, Jhi, &ll, &uu);
    if (lvec) std::copy(ll.begin(), ll.end(), lvec);
    if (uvec) std::copy(uu.begin(), uu.end(), uvec);
    return BBM_OK;
}

int32_t bbm_dist_sky_mm_layout(int64_t NrA, int64_t NcA, int64_t NcB, const int64_t* cA, const int64_t* cB, int64_t lA, int64_t uA,
                               int64_t lB, int64_t uB, int32_t nranks, int32_t rank, const int64_t* Kbounds, int64_t* ranges6,
                               int64_t* send_blocks, int64_t* recv_blocks) {
    MmLayout L;
    if ((NcA > 0 && !cA) || (NcB > 0 && !cB)) return BBM_ERR_ARG;
    int32_t rc = mm_layout(NrA, NcA, NcB, cA, cB, lA, uA, lB, uB, nranks, rank, Kbounds, &L);
    if (rc != BBM_OK) return rc;
    if (ranges6) { ranges6[0] = L.K0; ranges6[1] = L.K1; ranges6[2] = L.S0; ranges6[3] = L.S1; ranges6[4] = L.O0; ranges6[5] = L.O1; }
    for (int q = 0; q < nranks; ++q) {
        if (send_blocks) send_blocks[q] = (i64)L.sends[q].size();
        if (recv_blocks) recv_blocks[q] = (i64)L.recvs[q].size();
    }
    return BBM_OK;
}

int32_t bbm_dist_sky_plan_create(bbm_comm_t cm, int64_t Nr, int64_t Nc, const int64_t* r, const int64_t* c, int64_t l, int64_t u,
                                 const int64_t* Kbounds, bbm_dist_sky_plan_t* out) {
    if (!bbm_comm_valid(cm)) return BBM_ERR_HANDLE;
    bbm_handle_t h = cm->h;
    if (!out) return bbm_fail(h, BBM_ERR_ARG, "null out pointer");
    *out = nullptr;
    if (Nr < 0 || Nc < 0 || (Nr > 0 && !r) || (Nc > 0 && !c)) return bbm_fail(h, BBM_ERR_ARG, "bad block axes");
    std::vector<i64> kb(cm->nranks + 1);
    if (Kbounds) std::copy(Kbounds, Kbounds + cm->nranks + 1, kb.begin());
    else {
        int32_t rc = bbm_bbb_partition_rows(Nr, r, cm->nranks, kb.data());
        if (rc != BBM_OK) return bbm_fail(h, rc, "bad block sizes");
    }
    bbm_dist_sky_plan* pl = new (std::nothrow) bbm_dist_sky_plan();
    if (!pl) return bbm_fail(h, BBM_ERR_ALLOC, "out of host memory");
    pl->comm = cm; pl->h = h;
    int32_t rc = bbm_dist_layout_impl(Nr, Nc, r, c, l, u, cm->nranks, cm->rank, kb.data(), &pl->lay, &pl->sends, &pl->recvs);
    if (rc != BBM_OK) { delete pl; return bbm_fail(h, rc, "bad partition"); }
    const bbm_dist_layout& L = pl->lay;
    rowrange_structure(Nr, Nc, l, u, L.K0, L.K1, &pl->Jlo, &pl->Jhi, &pl->lloc, &pl->uloc);
    std::vector<i64> C(Nc + 1, 0);
    for (i64 J = 0; J < Nc; ++J) C[J + 1] = C[J] + c[J];
    pl->x_off = pl->Jhi > pl->Jlo ? C[pl->Jlo] - L.col_offset : 0;
    pl->n_cols = C[pl->Jhi] - C[pl->Jlo];
    rc = bbm_sky_desc_create(h, L.K1 - L.K0, pl->Jhi - pl->Jlo, r + L.K0, c + pl->Jlo, pl->lloc.data(), pl->uloc.data(), &pl->desc);
    if (rc != BBM_OK) { delete pl; return rc; }
    bbm_sky_desc_info(pl->desc, nullptr, nullptr, &pl->numel, nullptr);
    if (pl->x_off < 0 || pl->x_off + pl->n_cols > L.n_local) { bbm_dist_sky_plan_destroy(pl); return bbm_fail(h, BBM_ERR_UNSUPPORTED, "touched columns leave the local x range"); }
    *out = pl;
    return BBM_OK;
}

int32_t bbm_dist_sky_plan_destroy(bbm_dist_sky_plan_t pl) {
    if (!pl || pl->magic != 0xD157517Au) return BBM_ERR_HANDLE;
    if (pl->desc) bbm_sky_desc_destroy(pl->desc);
    pl->magic = 0;
    delete pl;
    return BBM_OK;
}

int32_t bbm_dist_sky_plan_layout(bbm_dist_sky_plan_t pl, bbm_dist_layout* out, int64_t* Jlo, int64_t* Jhi, int64_t* numentries_local,
                                 int64_t* x_offset) {
    if (!plan_valid(pl)) return BBM_ERR_HANDLE;
    if (out) *out = pl->lay;
    if (Jlo) *Jlo = pl->Jlo;
    if (Jhi) *Jhi = pl->Jhi;
    if (numentries_local) *numentries_local = pl->numel;
    if (x_offset) *x_offset = pl->x_off;
    return BBM_OK;
}

int32_t bbm_dist_sky_plan_bandwidths(bbm_dist_sky_plan_t pl, int64_t* lvec, int64_t* uvec) {
    if (!plan_valid(pl)) return BBM_ERR_HANDLE;
    if (lvec) std::copy(pl->lloc.begin(), pl->lloc.end(), lvec);
    if (uvec) std::copy(pl->uloc.begin(), pl->uloc.end(), uvec);
    return BBM_OK;
}

int32_t bbm_dist_sky_mul_f64(bbm_dist_sky_plan_t pl, double alpha, const double* d_data_local, double* d_x_local, int64_t nrhs, double beta,
                             double* d_y_own) { return dist_sky_mul_impl<double>(pl, alpha, d_data_local, d_x_local, nrhs, beta, d_y_own); }
int32_t bbm_dist_sky_mul_f32(bbm_dist_sky_plan_t pl, float alpha, const float* d_data_local, float* d_x_local, int64_t nrhs, float beta,
                             float* d_y_own) { return dist_sky_mul_impl<float>(pl, alpha, d_data_local, d_x_local, nrhs, beta, d_y_own); }

int32_t bbm_dist_sky_mm_plan_create(bbm_comm_t cm, int64_t NrA, int64_t NcA, int64_t NcB, const int64_t* rA, const int64_t* cA,
                                    const int64_t* cB, int64_t lA, int64_t uA, int64_t lB, int64_t uB, const int64_t* Kbounds,
                                    bbm_dist_sky_mm_plan_t* out) {
    if (!bbm_comm_valid(cm)) return BBM_ERR_HANDLE;
    bbm_handle_t h = cm->h;
    if (!out) return bbm_fail(h, BBM_ERR_ARG, "null out pointer");
    *out = nullptr;
    if (NrA < 0 || NcA < 0 || NcB < 0 || (NrA > 0 && !rA) || (NcA > 0 && !cA) || (NcB > 0 && !cB)) return bbm_fail(h, BBM_ERR_ARG, "bad block axes");
    std::vector<i64> kb(cm->nranks + 1);
    if (Kbounds) std::copy(Kbounds, Kbounds + cm->nranks + 1, kb.begin());
    else {
        int32_t rc = bbm_bbb_partition_rows(NrA, rA, cm->nranks, kb.data());
        if (rc != BBM_OK) return bbm_fail(h, rc, "bad block sizes");
    }
    bbm_dist_sky_mm_plan* pl = new (std::nothrow) bbm_dist_sky_mm_plan();
    if (!pl) return bbm_fail(h, BBM_ERR_ALLOC, "out of host memory");
    pl->comm = cm; pl->h = h;
    int32_t rc = mm_layout(NrA, NcA, NcB, cA, cB, lA, uA, lB, uB, cm->nranks, cm->rank, kb.data(), &pl->lay);
    if (rc != BBM_OK) { delete pl; return bbm_fail(h, rc, "bad partition"); }
    const MmLayout& L = pl->lay;
    // A_loc: rows [K0,K1) of A; B_loc: rows [S0,S1) of B; C_loc: rows [K0,K1) of C = A*B with bandwidths (lA+lB, uA+uB)
    const i64 rlo[3] = {L.K0, L.S0, L.K0}, rhi[3] = {L.K1, L.S1, L.K1};
    const i64 nr_[3] = {NrA, NcA, NrA}, nc_[3] = {NcA, NcB, NcB};
    const i64 ll[3] = {lA, lB, lA + lB}, uu[3] = {uA, uB, uA + uB};
    const i64* rws[3] = {rA, cA, rA};
    const i64* cls[3] = {cA, cB, cB};
    for (int t = 0; t < 3; ++t) {
        rowrange_structure(nr_[t], nc_[t], ll[t], uu[t], rlo[t], rhi[t], &pl->Jlo[t], &pl->Jhi[t], &pl->lloc[t], &pl->uloc[t]);
        rc = bbm_sky_desc_create(h, rhi[t] - rlo[t], pl->Jhi[t] - pl->Jlo[t], rws[t] + rlo[t], cls[t] + pl->Jlo[t], pl->lloc[t].data(),
                                 pl->uloc[t].data(), &pl->desc[t]);
        if (rc != BBM_OK) { bbm_dist_sky_mm_plan_destroy(pl); return rc; }
        bbm_sky_desc_info(pl->desc[t], nullptr, nullptr, &pl->numel[t], nullptr);
    }
    // the local product needs B_loc and C_loc over the same block columns; with non-negative bandwidths and a B at
    // least as wide as its rows reach they are ([K0-lA-lB, K1+uA+uB) clipped) -- anything else keeps the host path
    if ((pl->Jlo[2] != pl->Jlo[1] || pl->Jhi[2] != pl->Jhi[1]) && pl->numel[2] > 0) {
        bbm_dist_sky_mm_plan_destroy(pl);
        return bbm_fail(h, BBM_ERR_UNSUPPORTED, "the stored columns of B's and C's row ranges differ (degenerate bandwidths)");
    }
    // pack tables: blocks of B_loc in the (N,J) order both sides agree on
    SkyDescView bv;
    if (!bbm_sky_desc_view(pl->desc[1], &bv)) { bbm_dist_sky_mm_plan_destroy(pl); return bbm_fail(h, BBM_ERR_CUDA, "descriptor view"); }
    auto make_peer = [&](int q, const std::vector<BlockRef>& blocks, bbm_dist_sky_mm_plan::Peer* pr) -> int32_t {
        pr->q = q;
        std::vector<PackEnt> ents;
        i64 off = 0;
        for (const BlockRef& b : blocks) {
            const i64 Kp = b.N - L.S0, Jp = b.J - pl->Jlo[1];
            if (Kp < 0 || Kp >= bv.Nr || Jp < 0 || Jp >= bv.Nc || Kp < bv.klo[Jp] || Kp > bv.khi[Jp])
                return bbm_fail(h, BBM_ERR_BAND, "halo block (%lld,%lld) is not stored locally", (long long)b.N, (long long)b.J);
            PackEnt e{bv.off[Jp] + bv.R[Kp] - bv.R[bv.klo[Jp]], bv.S[Jp], off, (int32_t)bv.r[Kp], (int32_t)bv.c[Jp]};
            off += (i64)e.rows * e.cols;
            pr->maxblk = std::max<i64>(pr->maxblk, (i64)e.rows * e.cols);
            ents.push_back(e);
        }
        pr->count = off; pr->nents = (int)ents.size();
        if (!ents.empty()) {
            cudaError_t e = cudaMalloc(&pr->d_ents, ents.size() * sizeof(PackEnt));
            if (e == cudaSuccess) e = cudaMemcpy(pr->d_ents, ents.data(), ents.size() * sizeof(PackEnt), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { cudaGetLastError(); return bbm_fail(h, BBM_ERR_CUDA, "pack table upload: %s", cudaGetErrorString(e)); }
        }
        return BBM_OK;
    };
    cudaSetDevice(h->device);
    for (int q = 0; q < cm->nranks && rc == BBM_OK; ++q) {
        if (!L.sends[q].empty()) { pl->psend.emplace_back(); rc = make_peer(q, L.sends[q], &pl->psend.back()); }
        if (rc == BBM_OK && !L.recvs[q].empty()) { pl->precv.emplace_back(); rc = make_peer(q, L.recvs[q], &pl->precv.back()); }
    }
    if (rc != BBM_OK) { bbm_dist_sky_mm_plan_destroy(pl); return rc; }
    *out = pl;
    return BBM_OK;
}

int32_t bbm_dist_sky_mm_plan_destroy(bbm_dist_sky_mm_plan_t pl) {
    if (!pl || pl->magic != 0xD1575A33u) return BBM_ERR_HANDLE;
    if (bbm_valid(pl->h)) cudaSetDevice(pl->h->device);
    for (int t = 0; t < 3; ++t)
        if (pl->desc[t]) bbm_sky_desc_destroy(pl->desc[t]);
    for (auto* v : {&pl->psend, &pl->precv})
        for (auto& pr : *v) {
            if (pr.d_ents) cudaFree(pr.d_ents);
            if (pr.buf) cudaFree(pr.buf);
        }
    pl->magic = 0;
    delete pl;
    return BBM_OK;
}

int32_t bbm_dist_sky_mm_plan_info(bbm_dist_sky_mm_plan_t pl, int64_t* ranges6, int64_t* Jlo3, int64_t* Jhi3, int64_t* numentries3) {
    if (!mm_valid(pl)) return BBM_ERR_HANDLE;
    const MmLayout& L = pl->lay;
    if (ranges6) { ranges6[0] = L.K0; ranges6[1] = L.K1; ranges6[2] = L.S0; ranges6[3] = L.S1; ranges6[4] = L.O0; ranges6[5] = L.O1; }
    for (int t = 0; t < 3; ++t) {
        if (Jlo3) Jlo3[t] = pl->Jlo[t];
        if (Jhi3) Jhi3[t] = pl->Jhi[t];
        if (numentries3) numentries3[t] = pl->numel[t];
    }
    return BBM_OK;
}

int32_t bbm_dist_sky_mm_plan_bandwidths(bbm_dist_sky_mm_plan_t pl, int32_t which, int64_t* lvec, int64_t* uvec) {
    if (!mm_valid(pl)) return BBM_ERR_HANDLE;
    if (which < 0 || which > 2) return bbm_fail(pl->h, BBM_ERR_ARG, "which must be 0 (A), 1 (B) or 2 (C)");
    if (lvec) std::copy(pl->lloc[which].begin(), pl->lloc[which].end(), lvec);
    if (uvec) std::copy(pl->uloc[which].begin(), pl->uloc[which].end(), uvec);
    return BBM_OK;
}

int32_t bbm_dist_sky_matmul_f64(bbm_dist_sky_mm_plan_t pl, double alpha, const double* d_A_local, double* d_B_local, double beta,
                                double* d_C_local) { return dist_sky_matmul_impl<double>(pl, alpha, d_A_local, d_B_local, beta, d_C_local); }
int32_t bbm_dist_sky_matmul_f32(bbm_dist_sky_mm_plan_t pl, float alpha, const float* d_A_local, float* d_B_local, float beta,
                                float* d_C_local) { return dist_sky_matmul_impl<float>(pl, alpha, d_A_local, d_B_local, beta, d_C_local); }

}  // extern "C"
