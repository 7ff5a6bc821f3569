// Internal declarations shared by the multi-GPU translation units (dist.cu, distsky.cu): the NCCL entry points
// bound at run time and the communicator object.
#pragma once
#include <string>

#include "bbm_internal.h"

// ---- NCCL, bound at run time (libnccl.so.2 is only needed once a communicator is created) -------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess = 0 };
enum { ncclChar = 0 };

struct BbmNccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string err;
};
BbmNccl& bbm_nccl();
bool bbm_nccl_load();

struct bbm_comm {
    uint32_t magic = 0xC0221234u;
    bbm_handle_t h = nullptr;
    int nranks = 1, rank = 0;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;   // exchange stream
    cudaEvent_t ev_ready = nullptr;  // x ready on the compute stream
    cudaEvent_t ev_halo = nullptr;   // halo landed on the exchange stream
};
inline bool bbm_comm_valid(bbm_comm* c) { return c && c->magic == 0xC0221234u && bbm_valid(c->h); }

// host-only layout of one rank of a row-block partition (dist.cu); sends / recvs receive (peer, local column
// offset, element count) triples
int32_t bbm_dist_layout_impl(i64 Nr, i64 Nc, const i64* r, const i64* c, i64 l, i64 u, int nranks, int rank, const i64* Kb,
                             bbm_dist_layout* out, std::vector<i64>* sends, std::vector<i64>* recvs);
