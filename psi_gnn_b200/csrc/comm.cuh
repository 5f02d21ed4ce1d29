// comm.cuh — NCCL plumbing for the mesh-partitioned solve (one large mesh split by node ranges over the ranks).
//
// NCCL is resolved at run time with dlopen (the copy torch has already loaded is reused), so the library has no link-time
// dependency on it and loads on machines without NCCL.  All collectives are enqueued on the caller's stream: the solve
// stays a device-resident loop, no host round trip is added.
#pragma once
#include "common.cuh"
#include <dlfcn.h>
#include <vector>

// the few NCCL types we need (ABI-stable across NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;                     // ncclSuccess = 0
enum { psiNcclFloat32 = 7, psiNcclFloat64 = 8, psiNcclSum = 0 };   // ncclDataType_t / ncclRedOp_t values of nccl.h

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); if (h) break; }   // torch's copy, if loaded
    if (!h) for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) return nullptr;
    api.handle = h;
#define PSI_NCCL_SYM(field, sym) *(void**)(&api.field) = dlsym(h, sym); if (!api.field) { api.handle = nullptr; return nullptr; }
    PSI_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    PSI_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    PSI_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    PSI_NCCL_SYM(AllReduce, "ncclAllReduce")
    PSI_NCCL_SYM(Send, "ncclSend")
    PSI_NCCL_SYM(Recv, "ncclRecv")
    PSI_NCCL_SYM(GroupStart, "ncclGroupStart")
    PSI_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    PSI_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef PSI_NCCL_SYM
    return &api;
}

#define PSI_NCCL(call)                                                                      \
    do {                                                                                    \
        ncclResult_t r__ = (call);                                                          \
        if (r__ != 0) {                                                                     \
            g_psi_err = std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + #call + " -> " + \
                        (nccl_api() ? nccl_api()->GetErrorString(r__) : "nccl unavailable");  \
            return -1;                                                                      \
        }                                                                                   \
    } while (0)

struct psi_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

// Node-range partition of one mesh attached to a graph handle.  Local numbering: owned nodes [0, n_owned), then the
// ghost nodes grouped by owning peer in the order of `peers`.
struct Partition {
    psi_comm* comm = nullptr;
    int64_t n_owned = 0;
    std::vector<int> peers;
    std::vector<int64_t> send_count, recv_count, send_off, recv_off;   // in nodes
    int32_t* send_index = nullptr;          // device: concatenated local indices of owned rows to send, per peer
    float* send_buf = nullptr;              // device: [Σ send_count, 20] staging (10 floats per row for h, 20 for S̄)
    int64_t total_send = 0, total_recv = 0;
};

// gather rows of `width` floats (width even) into the contiguous staging buffer
__global__ void k_halo_pack(int64_t rows, int width, const int32_t* __restrict__ index, const float* __restrict__ src, float* __restrict__ dst,
                            const int* __restrict__ done) {
    if (done != nullptr && *done) return;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int per = width / 2;
    if (i >= rows * per) return;
    const int64_t r = i / per;
    const int c = (int)(i % per);
    reinterpret_cast<float2*>(dst)[i] = reinterpret_cast<const float2*>(src + (int64_t)index[r] * width)[c];
}

// exchange the ghost rows of `vec` ([n_loc, width] rows): owned rows listed in send_index go to the peers, the ghost segment is
// received in place.  Stream-ordered.
static int halo_exchange(Partition* P, float* vec, int width, const int* done, cudaStream_t st) {
    if (P == nullptr || P->comm == nullptr || P->peers.empty()) return 0;
    NcclApi* api = nccl_api();
    if (!api) PSI_FAIL("NCCL is not available");
    if (P->total_send > 0) {
        const int64_t work = P->total_send * (width / 2);
        k_halo_pack<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(P->total_send, width, P->send_index, vec, P->send_buf, done);
        PSI_CK_LAUNCH();
    }
    PSI_NCCL(api->GroupStart());
    for (size_t i = 0; i < P->peers.size(); ++i) {
        if (P->send_count[i] > 0)
            PSI_NCCL(api->Send(P->send_buf + P->send_off[i] * width, (size_t)P->send_count[i] * width, psiNcclFloat32, P->peers[i], P->comm->comm, st));
        if (P->recv_count[i] > 0)
            PSI_NCCL(api->Recv(vec + (P->n_owned + P->recv_off[i]) * width, (size_t)P->recv_count[i] * width, psiNcclFloat32, P->peers[i],
                               P->comm->comm, st));
    }
    PSI_NCCL(api->GroupEnd());
    return 0;
}

static int allreduce_f64(psi_comm* c, double* buf, size_t count, cudaStream_t st) {
    if (c == nullptr || c->world == 1) return 0;
    NcclApi* api = nccl_api();
    if (!api) PSI_FAIL("NCCL is not available");
    PSI_NCCL(api->AllReduce(buf, buf, count, psiNcclFloat64, psiNcclSum, c->comm, st));
    return 0;
}
