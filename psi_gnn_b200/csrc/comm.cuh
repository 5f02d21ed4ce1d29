// comm.cuh — NCCL plumbing for the mesh-partitioned solve (one large mesh split by node ranges over the ranks).
//
// NCCL is resolved at run time with dlopen (the copy torch has already loaded is reused), so the library has no link-time
// dependency on it and loads on machines without NCCL.  All collectives are enqueued on the caller's stream: the solve
// stays a device-resident loop, no host round trip is added.
#pragma once
#include "common.cuh"
#include "broyden.cuh"
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

// ---- peer-mapped mailboxes (NVLink/NVSwitch P2P, CUDA IPC): the device-initiated exchange of the partitioned solve --------------
// Every rank owns ONE cudaMalloc'd block that all other ranks map with cudaIpcOpenMemHandle:
//   [0, 4096)                      MailHeader: sequence flags written by the peers (monotone 64-bit counters, never reset)
//   [4096, xg_off)                 reduce slots  double red[2][world][PSI_RED_MAX]    (parity of the step, writer's rank)
//   [xg_off, sg_off)               ghost rows of the iterate   float xg[total_recv][PSI_QPITCH]     written by their owners
//   [sg_off, end)                  ghost rows of S̄ (VJP)       float sg[2][total_recv][PSI_QPITCH]
// A producer stores rows straight into the consumer's block, fences (__threadfence_system) and then publishes the step's sequence
// number in the consumer's header; the consumer spins on its own header (bounded: a time-out marks the solve as failed instead of
// hanging the GPU).  No NCCL call, no host round trip, no packing kernel on the critical path of a Broyden step.
#define PSI_MAX_WORLD 16
#define PSI_RED_MAX 8192                    // doubles per reduce slot: 3·threshold + 8 must fit
#define PSI_MAIL_RED_OFF 4096
#define PSI_SPIN_TIMEOUT_NS 4000000000ull   // 4 s
struct MailHeader {
    unsigned long long halo_seq[PSI_MAX_WORLD];      // written by the producer: its rows of exchange `seq` have landed here
    unsigned long long sb_seq[PSI_MAX_WORLD];
    unsigned long long red_seq[2][PSI_MAX_WORLD];
    unsigned long long halo_ack[PSI_MAX_WORLD];      // written by the consumer into the PRODUCER's header: rows of exchange `seq` consumed
    unsigned long long sb_ack[PSI_MAX_WORLD];
};
static inline size_t mail_xg_off(int world) { return PSI_MAIL_RED_OFF + (size_t)2 * world * PSI_RED_MAX * sizeof(double); }
static inline size_t mail_sg_off(int world, int64_t total_recv) { return mail_xg_off(world) + (size_t)total_recv * PSI_QPITCH * sizeof(float); }
static inline size_t mail_bytes(int world, int64_t total_recv) { return mail_sg_off(world, total_recv) + (size_t)2 * total_recv * PSI_QPITCH * sizeof(float) + 256; }

struct PeerDev {                            // one neighbour of the halo exchange (device-resident array)
    float* xg;                              // where my rows land in the peer's xg area (offset by my position in its ghost order)
    float* sg0; float* sg1;                 // … and in the two planes of its S̄ ghost area
    unsigned long long* halo_flag;          // &peer_header->halo_seq[my_rank]
    unsigned long long* sb_flag;            // &peer_header->sb_seq[my_rank]
    unsigned long long* halo_ack;           // &peer_header->halo_ack[my_rank]: where I acknowledge the rows the peer sent me
    unsigned long long* sb_ack;
    int send_off, send_count;               // my rows for this peer: send_index[send_off .. send_off + send_count)
    int rank, recv_off, recv_count;         // the peer's rank; its rows in MY ghost order
};
struct PartDev {                            // passed by value to the exchange kernels
    int n_peers, rank, world, pad_;
    const PeerDev* peers;                   // [n_peers]
    const int32_t* send_index;
    MailHeader* hdr;                        // my own header
    float* xg;                              // my own xg / sg areas
    float* sg;
    int64_t total_recv, total_send, n_owned, N;
    double* red_dst[PSI_MAX_WORLD];         // rank r's reduce area (red[0][rank of the writer = me] is at + rank·PSI_RED_MAX)
    unsigned long long* red_flag_dst[PSI_MAX_WORLD];   // &header_r->red_seq[0][me]  (parity p at + p·PSI_MAX_WORLD)
    unsigned int* counter;                  // last-block counters (device scratch, 4 of them)
    int* error;                             // set to 1 on a spin time-out
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spin until *flag ≥ seq; false on time-out
__device__ __forceinline__ bool spin_until(const unsigned long long* flag, unsigned long long seq) {
    if (ld_acquire_sys(flag) >= seq) return true;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flag) < seq) {
        __nanosleep(64);
        if (global_ns() - t0 > PSI_SPIN_TIMEOUT_NS) return false;
    }
    return true;
}

// Producer side of the halo exchange: every send row goes straight into its consumer's block (remote 16-byte stores over NVLink),
// the LAST block to finish publishes the sequence number at every neighbour.  plane_src: one [rows, pitch_src] array (the iterate:
// pitch 10) or the two S̄ planes (pitch PSI_QPITCH, second plane at + N·PSI_QPITCH).
template <int WHICH /*0 iterate rows → xg, 1 S̄ rows → sg*/>
__global__ void __launch_bounds__(128) k_halo_put(PartDev P, const float* __restrict__ src, unsigned long long seq, unsigned long long prev_seq,
                                                   int* done) {
    if (done != nullptr && *done) return;
    __shared__ int s_last, s_ok;
    // flow control: a neighbour's landing zone may be overwritten only after it has consumed the previous exchange (it says so in MY
    // header).  Inside a solve the all-reduce of every step already orders this; the first evaluations of a solve and stand-alone
    // exchanges have no such barrier.
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (prev_seq != 0 && threadIdx.x < P.n_peers) {
        const unsigned long long* a = (WHICH == 0 ? P.hdr->halo_ack : P.hdr->sb_ack) + P.peers[threadIdx.x].rank;
        if (!spin_until(a, prev_seq)) { s_ok = 0; *P.error = 1; if (done != nullptr) *done = 1; }
    }
    __syncthreads();
    if (!s_ok) return;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < P.total_send) {
        int pi = 0;
        while (pi + 1 < P.n_peers && i >= P.peers[pi + 1].send_off) ++pi;
        const PeerDev pd = P.peers[pi];
        const int64_t k = i - pd.send_off;
        const int64_t row = P.send_index[i];
        if (WHICH == 0) {
            const float2* s2 = reinterpret_cast<const float2*>(src + row * PSI_D);
            const float2 a = s2[0], b = s2[1], c = s2[2], d = s2[3], e = s2[4];
            float4* dst = reinterpret_cast<float4*>(pd.xg + k * PSI_QPITCH);
            dst[0] = make_float4(a.x, a.y, b.x, b.y);
            dst[1] = make_float4(c.x, c.y, d.x, d.y);
            dst[2] = make_float4(e.x, e.y, 0.f, 0.f);
        } else {
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                const float4* s4 = reinterpret_cast<const float4*>(src + ((int64_t)w * P.N + row) * PSI_QPITCH);
                float4* dst = reinterpret_cast<float4*>((w == 0 ? pd.sg0 : pd.sg1) + k * PSI_QPITCH);
                dst[0] = s4[0]; dst[1] = s4[1]; dst[2] = s4[2];
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&P.counter[WHICH], 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x < P.n_peers) st_release_sys(WHICH == 0 ? P.peers[threadIdx.x].halo_flag : P.peers[threadIdx.x].sb_flag, seq);
        if (threadIdx.x == 0) P.counter[WHICH] = 0;
    }
}

// Consumer side: wait for every neighbour's sequence number, then move the landed rows into the ghost segment of the local array
// (iterate: [N, 10] rows; S̄: the two [N, PSI_QPITCH] planes).  Every block waits on its own (the flags are monotone).
template <int WHICH>
__global__ void __launch_bounds__(128) k_halo_get(PartDev P, float* __restrict__ dst, unsigned long long seq, int* done) {
    if (done != nullptr && *done) return;
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (threadIdx.x < P.n_peers) {
        const unsigned long long* f = (WHICH == 0 ? P.hdr->halo_seq : P.hdr->sb_seq) + P.peers[threadIdx.x].rank;
        if (!spin_until(f, seq)) { s_ok = 0; *P.error = 1; if (done != nullptr) *done = 1; }   // a peer never arrived: abort the solve
    }
    __syncthreads();
    if (!s_ok) return;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < P.total_recv) {
        if (WHICH == 0) {
            const float4* s4 = reinterpret_cast<const float4*>(P.xg + i * PSI_QPITCH);
            const float4 a = __ldcg(s4), b = __ldcg(s4 + 1), c = __ldcg(s4 + 2);
            float2* d2 = reinterpret_cast<float2*>(dst + (P.n_owned + i) * PSI_D);
            d2[0] = make_float2(a.x, a.y); d2[1] = make_float2(a.z, a.w); d2[2] = make_float2(b.x, b.y); d2[3] = make_float2(b.z, b.w);
            d2[4] = make_float2(c.x, c.y);
        } else {
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                const float4* s4 = reinterpret_cast<const float4*>(P.sg + ((int64_t)w * P.total_recv + i) * PSI_QPITCH);
                float4* d4 = reinterpret_cast<float4*>(dst + ((int64_t)w * P.N + P.n_owned + i) * PSI_QPITCH);
                d4[0] = __ldcg(s4); d4[1] = __ldcg(s4 + 1); d4[2] = __ldcg(s4 + 2);
            }
        }
    }
    // the last block to finish tells every producer that its rows have been consumed
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&P.counter[3 + WHICH], 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        if (threadIdx.x < P.n_peers) st_release_sys(WHICH == 0 ? P.peers[threadIdx.x].halo_ack : P.peers[threadIdx.x].sb_ack, seq);
        if (threadIdx.x == 0) P.counter[3 + WHICH] = 0;
    }
}

// The all-reduce of a partitioned Broyden step fused behind the reduction kernel (replaces k_qn_fin1_local → ncclAllReduce →
// k_qn_fin1_global): every block reduces its rows of the partial matrix in fp64; the LAST block to finish writes the 3(n−1)+4
// local sums into the reduce slot of EVERY rank (remote stores), publishes the step's sequence number there, waits for the other
// ranks' numbers in its own header, adds the world's slots in rank order (identical on every rank: identical stop decisions) and
// evaluates the stopping rules.  One launch, one NVLink latency, no collective library on the path.
__global__ void __launch_bounds__(256)
k_qn_fin1_p2p(PartDev P, int nhist, const float* __restrict__ partial, int num_chunks, float* __restrict__ coef, int cap, double* __restrict__ dbuf,
              const float* __restrict__ norm_part, int norm_blocks, QnCtrl* __restrict__ ctrl, double* __restrict__ rel_trace,
              double* __restrict__ abs_trace, int step, double eps, double protect, int threshold, unsigned long long seq) {
    if (ctrl->done) return;
    __shared__ int s_last, s_ok;
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int cnt = 3 * nhist + 4;
    double* loc = dbuf + cnt + 8;                              // local sums (dbuf itself receives the global ones)
    for (int row = gwarp; row < nhist * 3 + 2; row += nwarps) {
        const float* p = partial + (int64_t)row * num_chunks;
        double s = 0.0;
        for (int i = lane; i < num_chunks; i += 32) s += (double)p[i];
        s = warp_sum_d(s);
        if (lane == 0) loc[row] = s;
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        double n1 = 0.0, n2 = 0.0;
        for (int i = lane; i < norm_blocks; i += 32) { n1 += (double)norm_part[i]; n2 += (double)norm_part[norm_blocks + i]; }
        n1 = warp_sum_d(n1);
        n2 = warp_sum_d(n2);
        if (lane == 0) { loc[3 * nhist + 2] = n1; loc[3 * nhist + 3] = n2; }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) { s_last = (atomicAdd(&P.counter[2], 1u) == gridDim.x - 1); s_ok = 1; }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int par = (int)(seq & 1ull);
    for (int r = 0; r < P.world; ++r) {
        double* dst = P.red_dst[r] + ((size_t)par * P.world + P.rank) * PSI_RED_MAX;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) dst[i] = __ldcg(loc + i);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < P.world) st_release_sys(P.red_flag_dst[threadIdx.x] + par * PSI_MAX_WORLD, seq);
    if (threadIdx.x < P.world) {
        if (!spin_until(&P.hdr->red_seq[par][threadIdx.x], seq)) { s_ok = 0; *P.error = 1; }
    }
    __syncthreads();
    if (threadIdx.x == 0) P.counter[2] = 0;
    if (!s_ok) {
        if (threadIdx.x == 0) { ctrl->done = 1; ctrl->stop_reason = 4; }
        return;
    }
    const double* mine = reinterpret_cast<const double*>(reinterpret_cast<const char*>(P.hdr) + PSI_MAIL_RED_OFF) + (size_t)par * P.world * PSI_RED_MAX;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < P.world; ++r) s += __ldcg(mine + (size_t)r * PSI_RED_MAX + i);
        dbuf[i] = s;
        if (i < nhist * 3) coef[(i % 3) * cap + i / 3] = (float)s;
    }
    __syncthreads();
    if (threadIdx.x < 32) qn_decide(threadIdx.x, dbuf[3 * nhist + 2], dbuf[3 * nhist + 3], ctrl, rel_trace, abs_trace, step, eps, protect, threshold);
}

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
    // peer-mapped mailboxes (psi_part_mail_open); without them the exchange falls back to the NCCL path above
    void* mail = nullptr;                   // my block (cudaMalloc)
    void* peer_mail[PSI_MAX_WORLD] = {};    // every rank's block as mapped here (self = mail)
    bool p2p = false;
    PartDev dev{};
    PeerDev* d_peers = nullptr;
    unsigned int* d_counter = nullptr;
    int* d_error = nullptr;
    // Sequence numbers of the three exchanges: (epoch << 24) + count.  Every solve starts a new epoch and restarts the counts, so the
    // numbers depend only on the (rank-independent) position of an exchange inside its solve — never on how many no-op steps a
    // rank's host happened to queue behind the stop.
    unsigned long long epoch = 0, cnt_halo = 0, cnt_sb = 0, cnt_red = 0;
    unsigned long long last_halo = 0, last_sb = 0;      // sequence numbers of the previous exchanges (what the peers must have acknowledged)
    void new_epoch() { ++epoch; cnt_halo = cnt_sb = cnt_red = 0; }
    unsigned long long next(unsigned long long& cnt) { return (epoch << 24) + (++cnt); }
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
