// psignn_b200.cu — C ABI of the B200-native PSI-GNN implicit message-passing solve (include/psignn_b200.h).
//
// Single translation unit: the kernels live in the .cuh files next to this one, this file holds the
// host side — handle lifetime, launch sequencing of the fused iteration, the device-resident solver loops.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC (psi_gnn_b200/build.py).
#include "../../include/psignn_b200.h"
#include "common.cuh"
#include "weights.cuh"
#include "graph.cuh"
#include "layer.cuh"
#include "vjp.cuh"
#include "broyden.cuh"
#include "anderson.cuh"
#include "qn_tma.cuh"
#include "comm.cuh"
#include "pgrad.cuh"
#include "baseline_bwd.cuh"

#include <algorithm>
#include <cstring>
#include <cmath>
#include <limits>
#include <vector>

thread_local std::string g_psi_err;

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline unsigned node_grid(int64_t N) { return (unsigned)((N + PSI_NODE_BLOCK - 1) / PSI_NODE_BLOCK); }

extern "C" int psi_version(void) { return 100; }
extern "C" const char* psi_last_error(void) { return g_psi_err.c_str(); }
extern "C" int psi_weights_floats(void) { return (int)((sizeof(LayerWeights) + sizeof(LayerWeightsT)) / sizeof(float)); }
#define PSI_WBLOB_FLOATS ((sizeof(LayerWeights) + sizeof(LayerWeightsT)) / sizeof(float))

// one packed block = LayerWeights followed by LayerWeightsT (the transposed copies for the packed-FMA kernels)
static int upload_block(const float* dev_blob, cudaStream_t st) {
    static_assert(sizeof(LayerBlock) == sizeof(LayerWeights) + sizeof(LayerWeightsT), "LayerBlock is the packed blob");
    PSI_CK(cudaMemcpyToSymbolAsync(cB, dev_blob, sizeof(LayerBlock), 0, cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int psi_weights_upload(const float* dev_blob, int n_floats, void* stream) {
    if (dev_blob == nullptr) PSI_FAIL("psi_weights_upload: null blob");
    if (n_floats != psi_weights_floats()) PSI_FAIL("psi_weights_upload: blob has the wrong number of floats");
    return upload_block(dev_blob, as_stream(stream));
}

// ================================================================================================
// graph
// ================================================================================================
static void part_release(Partition* P) {
    psi_free_async(P->send_index, nullptr);
    psi_free_async(P->send_buf, nullptr);
    if (P->mail != nullptr) {
        cudaDeviceSynchronize();                       // no kernel of ours may still be writing into a peer's block or reading ours
        for (int r = 0; r < PSI_MAX_WORLD; ++r)
            if (P->peer_mail[r] != nullptr && P->peer_mail[r] != P->mail) cudaIpcCloseMemHandle(P->peer_mail[r]);
        cudaFree(P->mail);
        if (P->d_peers) cudaFree(P->d_peers);
        if (P->d_counter) cudaFree(P->d_counter);
        if (P->d_error) cudaFree(P->d_error);
    }
}

static void graph_free(psi_graph* g) {
    void* ps[] = {g->p_recs_T, g->p_recs_F, g->p_recs_Ar, g->p_recs_Ac, g->p_off_T, g->p_off_F, g->p_off_Ar, g->p_off_Ac,
                  g->p_xm_T, g->p_xm_F, g->p_tag, g->p_prb, g->p_nrm, g->p_vjp, g->p_scratch, g->p_q, g->p_hstar};
    if (g->part != nullptr) {
        part_release(g->part);
        delete g->part;
    }
    for (void* p : ps)
        if (p) psi_free_async(p, nullptr);      // legacy default stream: ordered after the work torch's default stream has queued
    delete g;
}

extern "C" int psi_graph_create(psi_graph_t** out, int64_t num_nodes, int64_t nnz, const int64_t* dev_edge_index,
                                const float* dev_edge_attr, int attr_dim, const float* dev_a_ij, const float* dev_tags,
                                int tag_dim, const float* dev_prb, int prb_dim, const float* dev_normals, void* stream) {
    if (out == nullptr) PSI_FAIL("psi_graph_create: null out");
    *out = nullptr;
    if (num_nodes < 0 || nnz < 0) PSI_FAIL("psi_graph_create: negative size");
    if (3 * num_nodes >= (1ll << 31) - 64 || nnz >= (1ll << 31) - 64) PSI_FAIL("psi_graph_create: graph exceeds int32 indexing");
    if (attr_dim < 1 || attr_dim > 3) PSI_FAIL("psi_graph_create: attr_dim must be 1..3");
    if (prb_dim < 0 || prb_dim > 3) PSI_FAIL("psi_graph_create: prb_dim must be 0..3");
    if (dev_tags != nullptr && tag_dim != 1 && tag_dim != 3) PSI_FAIL("psi_graph_create: tag_dim must be 1 or 3");
    if (nnz > 0 && (dev_edge_index == nullptr || dev_edge_attr == nullptr)) PSI_FAIL("psi_graph_create: null edge arrays");
    if (num_nodes > 0 && prb_dim > 0 && dev_prb == nullptr) PSI_FAIL("psi_graph_create: null prb");
    cudaStream_t st = as_stream(stream);
    psi_graph* g = new psi_graph();
    g->N = num_nodes; g->nnz = nnz; g->attr_dim = attr_dim; g->prb_dim = prb_dim; g->tag_dim = tag_dim;
    const int64_t N1 = num_nodes > 0 ? num_nodes : 1;
    SellBuild bT, bF, bAr, bAc;
    if (build_sell(num_nodes, nnz, dev_edge_index, dev_edge_attr, attr_dim, nullptr, true, true, st, &bT)) { graph_free(g); return -1; }
    g->p_recs_T = bT.recs; g->p_off_T = bT.off; g->slots_T = bT.slots;
    if (build_sell(num_nodes, nnz, dev_edge_index, dev_edge_attr, attr_dim, nullptr, true, false, st, &bF)) { graph_free(g); return -1; }
    g->p_recs_F = bF.recs; g->p_off_F = bF.off; g->slots_F = bF.slots;
    g->E = bT.kept;
    if (dev_a_ij != nullptr) {
        if (build_sell(num_nodes, nnz, dev_edge_index, nullptr, 0, dev_a_ij, false, false, st, &bAr)) { graph_free(g); return -1; }
        g->p_recs_Ar = bAr.recs; g->p_off_Ar = bAr.off; g->slots_Ar = bAr.slots;
        if (build_sell(num_nodes, nnz, dev_edge_index, nullptr, 0, dev_a_ij, false, true, st, &bAc)) { graph_free(g); return -1; }
        g->p_recs_Ac = bAc.recs; g->p_off_Ac = bAc.off; g->slots_Ac = bAc.slots;
    }
    // node data
    unsigned long long* counts = nullptr;
    if (psi_malloc_async(&g->p_tag, N1, st) != cudaSuccess || psi_malloc_async((void**)&counts, 2 * sizeof(unsigned long long), st) != cudaSuccess) {
        graph_free(g); PSI_FAIL("psi_graph_create: out of device memory");
    }
    cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned long long), st);
    if (num_nodes > 0) {
        k_graph_tags<<<(unsigned)((num_nodes + 255) / 256), 256, 0, st>>>((int)num_nodes, dev_tags, tag_dim, (uint8_t*)g->p_tag, counts);
    }
    if (prb_dim > 0) {
        if (psi_malloc_async(&g->p_prb, N1 * prb_dim * sizeof(float), st) != cudaSuccess) { graph_free(g); PSI_FAIL("psi_graph_create: out of device memory"); }
        if (num_nodes > 0) cudaMemcpyAsync(g->p_prb, dev_prb, num_nodes * prb_dim * sizeof(float), cudaMemcpyDeviceToDevice, st);
    }
    if (dev_normals != nullptr) {
        if (psi_malloc_async(&g->p_nrm, N1 * 2 * sizeof(float), st) != cudaSuccess) { graph_free(g); PSI_FAIL("psi_graph_create: out of device memory"); }
        if (num_nodes > 0) cudaMemcpyAsync(g->p_nrm, dev_normals, num_nodes * 2 * sizeof(float), cudaMemcpyDeviceToDevice, st);
    }
    g->scratch_floats = 2 * (int64_t)node_grid(N1) + 8;
    if (psi_malloc_async((void**)&g->p_scratch, g->scratch_floats * sizeof(float), st) != cudaSuccess) { graph_free(g); PSI_FAIL("psi_graph_create: out of device memory"); }
    unsigned long long hc[2] = {0, 0};
    cudaMemcpyAsync(hc, counts, sizeof(hc), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    psi_free_async(counts, st);
    if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) {
        graph_free(g);
        PSI_FAIL(std::string("psi_graph_create: ") + cudaGetErrorString(e));
    }
    g->n_dir = (int64_t)hc[0]; g->n_neu = (int64_t)hc[1];
    GraphDev& D = g->dev;
    D.N = (int)num_nodes;
    D.n_compute = (int)num_nodes;
    D.num_slices = (int)((num_nodes + 31) / 32);
    D.prb_dim = prb_dim;
    D.T = SellDev{(const int4*)g->p_recs_T, nullptr, (const int64_t*)g->p_off_T, nullptr};
    D.F = SellDev{(const int4*)g->p_recs_F, nullptr, (const int64_t*)g->p_off_F, nullptr};
    D.Ar = SellDev{nullptr, (const int2*)g->p_recs_Ar, (const int64_t*)g->p_off_Ar, nullptr};
    D.Ac = SellDev{nullptr, (const int2*)g->p_recs_Ac, (const int64_t*)g->p_off_Ac, nullptr};
    D.tag = (const uint8_t*)g->p_tag;
    D.prb = (const float*)g->p_prb;
    D.nrm = (const float*)g->p_nrm;
    g->bytes = g->slots_T * 16 + g->slots_F * 16 + g->slots_Ar * 8 + g->slots_Ac * 8 + 4 * (D.num_slices + 1) * 8 + N1 +
               N1 * prb_dim * 4 + (dev_normals ? N1 * 8 : 0) + g->scratch_floats * 4;
    *out = g;
    return 0;
}

extern "C" int psi_graph_destroy(psi_graph_t* g) {
    if (g == nullptr) return 0;
    graph_free(g);
    return 0;
}

extern "C" int psi_graph_info(const psi_graph_t* g, int64_t info[8]) {
    if (g == nullptr || info == nullptr) PSI_FAIL("psi_graph_info: null argument");
    info[0] = g->N; info[1] = g->E; info[2] = g->nnz; info[3] = g->n_dir; info[4] = g->n_neu;
    info[5] = g->slots_T; info[6] = g->slots_F; info[7] = g->bytes;
    return 0;
}

// ================================================================================================
// communicator and mesh partition
// ================================================================================================
extern "C" int psi_comm_unique_id(char out[128]) {
    NcclApi* api = nccl_api();
    if (!api) PSI_FAIL("psi_comm_unique_id: NCCL is not available");
    ncclUniqueId id;
    PSI_NCCL(api->GetUniqueId(&id));
    memcpy(out, id.internal, 128);
    return 0;
}

extern "C" int psi_comm_create(psi_comm_t** out, int rank, int world, const char id[128]) {
    if (out == nullptr) PSI_FAIL("psi_comm_create: null out");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) PSI_FAIL("psi_comm_create: bad rank/world");
    psi_comm* c = new psi_comm();
    c->rank = rank; c->world = world;
    if (world > 1) {
        NcclApi* api = nccl_api();
        if (!api) { delete c; PSI_FAIL("psi_comm_create: NCCL is not available"); }
        ncclUniqueId uid;
        memcpy(uid.internal, id, 128);
        ncclResult_t r = api->CommInitRank(&c->comm, world, uid, rank);
        if (r != 0) { delete c; PSI_FAIL(std::string("psi_comm_create: ncclCommInitRank -> ") + api->GetErrorString(r)); }
    }
    *out = c;
    return 0;
}

extern "C" int psi_comm_destroy(psi_comm_t* c) {
    if (c == nullptr) return 0;
    if (c->comm != nullptr && nccl_api()) nccl_api()->CommDestroy(c->comm);
    delete c;
    return 0;
}

extern "C" int psi_graph_set_partition(psi_graph_t* g, psi_comm_t* comm, int64_t n_owned, int n_peers, const int32_t* peer_ranks,
                                       const int64_t* send_counts, const int64_t* recv_counts, const int32_t* dev_send_index, void* stream) {
    if (g == nullptr || comm == nullptr) PSI_FAIL("psi_graph_set_partition: null handle");
    if (n_owned < 0 || n_owned > g->N || n_peers < 0) PSI_FAIL("psi_graph_set_partition: bad sizes");
    cudaStream_t st = as_stream(stream);
    Partition* P = new Partition();
    P->comm = comm; P->n_owned = n_owned;
    for (int i = 0; i < n_peers; ++i) {
        P->peers.push_back(peer_ranks[i]);
        P->send_off.push_back(P->total_send); P->recv_off.push_back(P->total_recv);
        P->send_count.push_back(send_counts[i]); P->recv_count.push_back(recv_counts[i]);
        P->total_send += send_counts[i]; P->total_recv += recv_counts[i];
    }
    if (n_owned + P->total_recv != g->N) { delete P; PSI_FAIL("psi_graph_set_partition: owned + ghost rows must equal the graph's node count"); }
    if (P->total_send > 0) {
        if (dev_send_index == nullptr) { delete P; PSI_FAIL("psi_graph_set_partition: null send index"); }
        if (psi_malloc_async((void**)&P->send_index, P->total_send * sizeof(int32_t), st) != cudaSuccess ||
            psi_malloc_async((void**)&P->send_buf, P->total_send * 20 * sizeof(float), st) != cudaSuccess) { delete P; PSI_FAIL("psi_graph_set_partition: out of device memory"); }
        PSI_CK(cudaMemcpyAsync(P->send_index, dev_send_index, P->total_send * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    }
    if (g->part != nullptr) { part_release(g->part); delete g->part; }
    g->part = P;
    g->dev.n_compute = (int)n_owned;
    return 0;
}

// ---- peer-mapped mailboxes: create my block (returns its IPC handle), then map every rank's block ------------------------------
extern "C" int psi_part_mail_create(psi_graph_t* g, char handle_out[64], int64_t* total_recv_out) {
    if (g == nullptr || g->part == nullptr) PSI_FAIL("psi_part_mail_create: the graph has no partition");
    Partition* P = g->part;
    const int world = P->comm->world;
    if (world > PSI_MAX_WORLD) PSI_FAIL("psi_part_mail_create: world size exceeds PSI_MAX_WORLD");
    if (P->mail == nullptr) {
        const size_t bytes = mail_bytes(world, P->total_recv);
        PSI_CK(cudaMalloc(&P->mail, bytes));
        PSI_CK(cudaMemset(P->mail, 0, bytes));
        PSI_CK(cudaMalloc((void**)&P->d_counter, 8 * sizeof(unsigned int)));
        PSI_CK(cudaMemset(P->d_counter, 0, 8 * sizeof(unsigned int)));
        PSI_CK(cudaMalloc((void**)&P->d_error, sizeof(int)));
        PSI_CK(cudaMemset(P->d_error, 0, sizeof(int)));
        PSI_CK(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    PSI_CK(cudaIpcGetMemHandle(&h, P->mail));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle_out, &h, 64);
    if (total_recv_out) *total_recv_out = P->total_recv;
    return 0;
}

// handles: [world][64] IPC handles of every rank's block (own entry ignored); total_recvs: [world]; remote_off: [n_peers] = position
// of MY rows in the ghost order of neighbour i (rows)
extern "C" int psi_part_mail_open(psi_graph_t* g, const char* handles, const int64_t* total_recvs, const int64_t* remote_off) {
    if (g == nullptr || g->part == nullptr || g->part->mail == nullptr) PSI_FAIL("psi_part_mail_open: call psi_part_mail_create first");
    Partition* P = g->part;
    const int world = P->comm->world, rank = P->comm->rank;
    for (int r = 0; r < world; ++r) {
        if (r == rank) { P->peer_mail[r] = P->mail; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        PSI_CK(cudaIpcOpenMemHandle(&P->peer_mail[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    std::vector<PeerDev> pd(P->peers.size());
    for (size_t i = 0; i < P->peers.size(); ++i) {
        const int q = P->peers[i];
        char* base = (char*)P->peer_mail[q];
        MailHeader* hq = (MailHeader*)base;
        float* xg = (float*)(base + mail_xg_off(world));
        float* sg = (float*)(base + mail_sg_off(world, total_recvs[q]));
        pd[i].xg = xg + remote_off[i] * PSI_QPITCH;
        pd[i].sg0 = sg + remote_off[i] * PSI_QPITCH;
        pd[i].sg1 = sg + (total_recvs[q] + remote_off[i]) * PSI_QPITCH;
        pd[i].halo_flag = &hq->halo_seq[rank];
        pd[i].sb_flag = &hq->sb_seq[rank];
        pd[i].halo_ack = &hq->halo_ack[rank];
        pd[i].sb_ack = &hq->sb_ack[rank];
        pd[i].send_off = (int)P->send_off[i]; pd[i].send_count = (int)P->send_count[i];
        pd[i].rank = q; pd[i].recv_off = (int)P->recv_off[i]; pd[i].recv_count = (int)P->recv_count[i];
    }
    if (P->d_peers) cudaFree(P->d_peers);
    PSI_CK(cudaMalloc((void**)&P->d_peers, std::max<size_t>(1, pd.size()) * sizeof(PeerDev)));
    if (!pd.empty()) PSI_CK(cudaMemcpy(P->d_peers, pd.data(), pd.size() * sizeof(PeerDev), cudaMemcpyHostToDevice));
    PartDev& D = P->dev;
    D.n_peers = (int)pd.size(); D.rank = rank; D.world = world; D.peers = P->d_peers; D.send_index = P->send_index;
    D.hdr = (MailHeader*)P->mail;
    D.xg = (float*)((char*)P->mail + mail_xg_off(world));
    D.sg = (float*)((char*)P->mail + mail_sg_off(world, P->total_recv));
    D.total_recv = P->total_recv; D.total_send = P->total_send; D.n_owned = P->n_owned; D.N = g->N;
    for (int r = 0; r < world; ++r) {
        char* base = (char*)P->peer_mail[r];
        D.red_dst[r] = (double*)(base + PSI_MAIL_RED_OFF);
        D.red_flag_dst[r] = &((MailHeader*)base)->red_seq[0][rank];
    }
    D.counter = P->d_counter; D.error = P->d_error;
    if (D.n_peers > 32) PSI_FAIL("psi_part_mail_open: more than 32 neighbour ranks");
    P->p2p = true;
    return 0;
}

// 1 if a device-side exchange of this partition timed out (the peers did not arrive): the solve's results are invalid
extern "C" int psi_part_error(const psi_graph_t* g) {
    if (g == nullptr || g->part == nullptr || g->part->d_error == nullptr) return 0;
    int e = 0;
    cudaMemcpy(&e, g->part->d_error, sizeof(int), cudaMemcpyDeviceToHost);
    return e;
}

// ghost rows of the iterate (which = 0, [N,10] rows) or of the two S̄ planes (which = 1, [2][N][12]) from their owners
static int halo_refresh(psi_graph* g, float* vec, int which, const int* done, cudaStream_t st) {
    Partition* P = g->part;
    if (P == nullptr || P->comm == nullptr || P->peers.empty()) return 0;
    if (P->p2p) {
        PartDev D = P->dev;
        D.N = g->N;
        int* dn = const_cast<int*>(done);
        const unsigned pgrid = (unsigned)std::max<int64_t>(1, (P->total_send + 127) / 128);
        const unsigned ggrid = (unsigned)std::max<int64_t>(1, (P->total_recv + 127) / 128);
        if (which == 0) {
            const unsigned long long seq = P->next(P->cnt_halo);
            k_halo_put<0><<<pgrid, 128, 0, st>>>(D, vec, seq, P->last_halo, dn);
            k_halo_get<0><<<ggrid, 128, 0, st>>>(D, vec, seq, dn);
            P->last_halo = seq;
        } else {
            const unsigned long long seq = P->next(P->cnt_sb);
            k_halo_put<1><<<pgrid, 128, 0, st>>>(D, vec, seq, P->last_sb, dn);
            k_halo_get<1><<<ggrid, 128, 0, st>>>(D, vec, seq, dn);
            P->last_sb = seq;
        }
        PSI_CK_LAUNCH();
        return 0;
    }
    // fallback without mapped peer memory: pack + grouped ncclSend/ncclRecv
    if (which == 0) return halo_exchange(P, vec, PSI_D, done, st);
    if (halo_exchange(P, vec, PSI_QPITCH, done, st)) return -1;
    return halo_exchange(P, vec + g->N * PSI_QPITCH, PSI_QPITCH, done, st);
}

// refresh the ghost rows of a [N, width] array from their owners (no-op for an unpartitioned graph)
extern "C" int psi_halo_exchange(psi_graph_t* g, float* dev_vec, int width, void* stream) {
    if (g == nullptr) PSI_FAIL("psi_halo_exchange: null graph");
    if (width != 10 && width != 20 && width != 2) PSI_FAIL("psi_halo_exchange: width must be 2, 10 or 20");
    if (width == 10) return halo_refresh(g, dev_vec, 0, nullptr, as_stream(stream));
    return halo_exchange(g->part, dev_vec, width, nullptr, as_stream(stream));
}

static int check_kind(const psi_graph* g, int kind) {
    if (g == nullptr) PSI_FAIL("null graph handle");
    if (kind < 0 || kind > 4) PSI_FAIL("unknown layer kind");
    const bool mixed = (kind == PSI_KIND_MIXED || kind == PSI_KIND_DSGPS_MIXED);
    const int want_prb = (mixed || kind == PSI_KIND_DSS) ? 3 : 2;
    if (g->prb_dim != want_prb) PSI_FAIL("graph second-member width does not match the layer kind");
    if (mixed && g->p_nrm == nullptr) PSI_FAIL("mixed layer needs unit normals");
    if (mixed && g->tag_dim != 3) PSI_FAIL("mixed layer needs 3-column one-hot tags");
    if (kind == PSI_KIND_DSS && g->attr_dim != 1) PSI_FAIL("DSS layer needs 1 edge attribute");
    if (kind != PSI_KIND_DSS && g->attr_dim != 3) PSI_FAIL("layer needs 3 edge attributes");
    return 0;
}

// ================================================================================================
// layer / VJP / residual / encoder / decoder
// ================================================================================================
// scratch of the layer pre-pass: Q[w][node] = W1j_w·h[node], w < 3 (allocated on first use, lives with the handle)
static int graph_q(const psi_graph* cg, cudaStream_t st, float** q) {
    psi_graph* g = const_cast<psi_graph*>(cg);
    if (g->p_q == nullptr) {
        const int64_t N1 = g->N > 0 ? g->N : 1;
        PSI_CK(psi_malloc_async((void**)&g->p_q, (size_t)3 * N1 * PSI_QPITCH * sizeof(float), st));
        g->bytes += 3 * N1 * PSI_QPITCH * 4;
    }
    *q = g->p_q;
    return 0;
}

// grids of at most two CTAs per SM (C0 / C1 / C2-size batches): the operator is latency-bound, one launch less is worth more than the
// per-edge recomputation of W1j·h_j costs (layer.cuh walk_direct_h; bit-identical results)
static inline bool small_grid(unsigned grid) {
    static const bool off = getenv("PSI_NO_FUSED_PRE") != nullptr;              // A/B switch
    return !off && grid <= 2u * PSI_NUM_SMS_B200;
}

static inline bool fused_pre(const psi_graph* g, int kind) {
    const bool mixed = (kind == PSI_KIND_MIXED || kind == PSI_KIND_DSGPS_MIXED);
    return !mixed && g->part == nullptr && small_grid(node_grid(g->dev.n_compute));
}

// one application of the layer = pre-pass (per-source half of the first edge layer) + fused gather/update kernel: 2 launches
// (1 launch on small grids for the kinds without a Neumann edge MLP)
template <bool EPI>
static int launch_layer(const psi_graph* g, int kind, const float* h, const float* h0, float* out, SolverEpi E, cudaStream_t st) {
    if (g->dev.n_compute == 0) return 0;
    float* Q = nullptr;
    if (graph_q(g, st, &Q)) return -1;
    const unsigned grid = node_grid(g->dev.n_compute);
    const unsigned pre_grid = node_grid(g->N);           // every row that can be a message source (owned + ghost rows)
    const bool mixed = (kind == PSI_KIND_MIXED || kind == PSI_KIND_DSGPS_MIXED);
    if (fused_pre(g, kind)) {
        switch (kind) {
            case PSI_KIND_DIRICHLET: k_layer_forward<KIND_DIRICHLET, EPI, true><<<grid, PSI_NODE_BLOCK, 0, st>>>(g->dev, h, h0, Q, out, E); break;
            case PSI_KIND_DSS:       k_layer_forward<KIND_DSS, EPI, true><<<grid, PSI_NODE_BLOCK, 0, st>>>(g->dev, h, h0, Q, out, E); break;
            default:                 k_layer_forward<KIND_DSGPS, EPI, true><<<grid, PSI_NODE_BLOCK, 0, st>>>(g->dev, h, h0, Q, out, E); break;
        }
        PSI_CK_LAUNCH();
        return 0;
    }
    if (mixed) k_layer_pre<3><<<pre_grid, PSI_NODE_BLOCK, 0, st>>>((int)g->N, h, Q, E.done);
    else k_layer_pre<2><<<pre_grid, PSI_NODE_BLOCK, 0, st>>>((int)g->N, h, Q, E.done);
    switch (kind) {
        case PSI_KIND_DIRICHLET:   k_layer_forward<KIND_DIRICHLET, EPI><<<grid, PSI_NODE_BLOCK, 0, st>>>(g->dev, h, h0, Q, out, E); break;
        case PSI_KIND_MIXED:       k_layer_forward<KIND_MIXED, EPI><<<grid, PSI_NODE_BLOCK, 0, st>>>(g->dev, h, h0, Q, out, E); break;
        case PSI_KIND_DSS:         k_layer_forward<KIND_DSS, EPI><<<grid, PSI_NODE_BLOCK, 0, st>>>(g->dev, h, h0, Q, out, E); break;
        case PSI_KIND_DSGPS:       k_layer_forward<KIND_DSGPS, EPI><<<grid, PSI_NODE_BLOCK, 0, st>>>(g->dev, h, h0, Q, out, E); break;
        default:                   k_layer_forward<KIND_DSGPS_MIXED, EPI><<<grid, PSI_NODE_BLOCK, 0, st>>>(g->dev, h, h0, Q, out, E); break;
    }
    PSI_CK_LAUNCH();
    return 0;
}

extern "C" int psi_layer_forward(const psi_graph_t* g, int kind, const float* dev_h, const float* dev_h0, float* dev_out, void* stream) {
    if (check_kind(g, kind)) return -1;
    if (g->N > 0 && (dev_h == nullptr || dev_out == nullptr)) PSI_FAIL("psi_layer_forward: null pointer");
    if (g->N > 0 && kind != PSI_KIND_DSS && dev_h0 == nullptr) PSI_FAIL("psi_layer_forward: null h0");
    if (g->N > 0 && dev_h == dev_out) PSI_FAIL("psi_layer_forward: in-place application is not supported");
    return launch_layer<false>(g, kind, dev_h, dev_h0, dev_out, SolverEpi{nullptr, nullptr, nullptr, nullptr}, as_stream(stream));
}

// k unrolled applications of the layer (DSS: one weight block per layer; DSGPS: the same block k times) in one call:
// out = f_{k-1}(… f_1(f_0(h)) …).  dev_work is a scratch [N, d] buffer for the ping-pong.
extern "C" int psi_layers_unrolled(const psi_graph_t* g, int kind, const float* dev_blobs, int n_blobs, int n_layers, const float* dev_h,
                                   const float* dev_h0, float* dev_work, float* dev_out, void* stream) {
    if (check_kind(g, kind)) return -1;
    if (n_layers < 1 || (n_blobs != 1 && n_blobs != n_layers)) PSI_FAIL("psi_layers_unrolled: n_blobs must be 1 or n_layers");
    if (dev_blobs == nullptr) PSI_FAIL("psi_layers_unrolled: null weight blocks");
    if (g->N > 0 && (dev_h == nullptr || dev_out == nullptr || dev_work == nullptr)) PSI_FAIL("psi_layers_unrolled: null pointer");
    if (g->N > 0 && (dev_h == dev_out || dev_h == dev_work || dev_out == dev_work)) PSI_FAIL("psi_layers_unrolled: buffers must be distinct");
    cudaStream_t st = as_stream(stream);
    const SolverEpi noE{nullptr, nullptr, nullptr, nullptr};
    const size_t wf = PSI_WBLOB_FLOATS;
    const float* src = dev_h;
    for (int k = 0; k < n_layers; ++k) {
        if (k == 0 || n_blobs > 1)
            if (upload_block(dev_blobs + (size_t)(n_blobs > 1 ? k : 0) * wf, st)) return -1;
        // ping-pong so that the last layer lands in dev_out
        float* dst = ((n_layers - 1 - k) % 2 == 0) ? dev_out : dev_work;
        if (launch_layer<false>(g, kind, src, dev_h0, dst, noE, st)) return -1;
        src = dst;
    }
    return 0;
}

static int vjp_alloc(psi_graph* g, cudaStream_t st) {
    if (g->p_vjp != nullptr) return 0;
    const int64_t N = g->N > 0 ? g->N : 1;
    const int64_t floats = N * (10 + 1 + 1 + 10 + 30 + 1 + 2 * PSI_QPITCH + 10);
    PSI_CK(psi_malloc_async(&g->p_vjp, floats * sizeof(float), st));
    PSI_CK(psi_malloc_async(&g->p_xm_T, (g->slots_T > 0 ? g->slots_T : 1) * sizeof(int2), st));
    PSI_CK(psi_malloc_async(&g->p_xm_F, (g->slots_F > 0 ? g->slots_F : 1) * sizeof(int2), st));
    float* p = (float*)g->p_vjp;
    VjpCacheDev& C = g->vjp;
    C.Sb = p; p += N * 2 * PSI_QPITCH;          // first: its rows are read with 16-byte loads (the base is 256-byte aligned)
    C.rhat = p; p += N * 10;
    C.m = p; p += N * 10;
    C.cnt = p; p += N * 30;
    C.Dloc = p; p += N * 10;
    C.rstd = p; p += N;
    C.alpha = p; p += N;
    C.nmask = (uint32_t*)p; p += N;
    g->dev.T.xmask = (uint32_t*)g->p_xm_T;
    g->dev.F.xmask = (uint32_t*)g->p_xm_F;
    g->bytes += floats * 4 + (g->slots_T + g->slots_F) * 8;
    return 0;
}

extern "C" int psi_vjp_prepare(psi_graph_t* g, int kind, const float* dev_hstar, const float* dev_h0, void* stream) {
    (void)dev_h0;   // the Dirichlet rows of f do not depend on h: h0 never enters the Jacobian
    if (check_kind(g, kind)) return -1;
    if (kind != PSI_KIND_DIRICHLET && kind != PSI_KIND_MIXED) PSI_FAIL("psi_vjp_prepare: VJP exists for the PSI-GNN layers only");
    if (g->N > 0 && dev_hstar == nullptr) PSI_FAIL("psi_vjp_prepare: null pointer");
    cudaStream_t st = as_stream(stream);
    if (vjp_alloc(g, st)) return -1;
    if (g->part != nullptr && g->N > 0) {
        // mesh partition: the linearisation point needs the neighbours' rows of H* — take a private copy and refresh its ghost rows
        if (g->p_hstar == nullptr) {
            PSI_CK(psi_malloc_async((void**)&g->p_hstar, (size_t)g->N * PSI_D * sizeof(float), st));
            g->bytes += g->N * PSI_D * 4;
        }
        PSI_CK(cudaMemcpyAsync(g->p_hstar, dev_hstar, (size_t)g->N * PSI_D * sizeof(float), cudaMemcpyDeviceToDevice, st));
        if (halo_refresh(g, g->p_hstar, 0, nullptr, st)) return -1;
        dev_hstar = g->p_hstar;
    }
    if (g->N > 0) {
        PSI_CK(cudaMemsetAsync(g->p_xm_T, 0, (g->slots_T > 0 ? g->slots_T : 1) * sizeof(int2), st));
        PSI_CK(cudaMemsetAsync(g->p_xm_F, 0, (g->slots_F > 0 ? g->slots_F : 1) * sizeof(int2), st));
        if (kind == PSI_KIND_DIRICHLET) k_vjp_prepare<KIND_DIRICHLET><<<node_grid(g->N), PSI_NODE_BLOCK, 0, st>>>(g->dev, g->vjp, dev_hstar);
        else k_vjp_prepare<KIND_MIXED><<<node_grid(g->N), PSI_NODE_BLOCK, 0, st>>>(g->dev, g->vjp, dev_hstar);
        PSI_CK_LAUNCH();
    }
    g->vjp_ready = true;
    g->vjp_kind = kind;
    return 0;
}

template <bool EPI>
static int launch_vjp(psi_graph* g, int kind, const float* y, const float* grad, float* out, SolverEpi E, cudaStream_t st, float* acc_out = nullptr) {
    if (g->dev.n_compute == 0) return 0;
    const unsigned grid = node_grid(g->dev.n_compute);
    if (kind == PSI_KIND_DIRICHLET) k_vjp_phase_a<KIND_DIRICHLET><<<grid, PSI_NODE_BLOCK, 0, st>>>(g->dev, g->vjp, y, E.done);
    else k_vjp_phase_a<KIND_MIXED><<<grid, PSI_NODE_BLOCK, 0, st>>>(g->dev, g->vjp, y, E.done);
    // mesh partition: S̄ of the ghost rows comes from their owners between the two phases (y itself is needed on owned rows only)
    if (g->part != nullptr && halo_refresh(g, g->vjp.Sb, 1, E.done, st)) return -1;
    if (kind == PSI_KIND_DIRICHLET) k_vjp_phase_b<KIND_DIRICHLET, EPI><<<grid, PSI_NODE_BLOCK, 0, st>>>(g->dev, g->vjp, y, grad, out, E, acc_out);
    else k_vjp_phase_b<KIND_MIXED, EPI><<<grid, PSI_NODE_BLOCK, 0, st>>>(g->dev, g->vjp, y, grad, out, E, acc_out);
    PSI_CK_LAUNCH();
    return 0;
}

extern "C" int psi_vjp_apply(psi_graph_t* g, int kind, const float* dev_y, const float* dev_grad, float* dev_out, void* stream) {
    if (check_kind(g, kind)) return -1;
    if (!g->vjp_ready || g->vjp_kind != kind) PSI_FAIL("psi_vjp_apply: call psi_vjp_prepare first");
    if (g->N > 0 && (dev_y == nullptr || dev_out == nullptr)) PSI_FAIL("psi_vjp_apply: null pointer");
    return launch_vjp<false>(g, kind, dev_y, dev_grad, dev_out, SolverEpi{nullptr, nullptr, nullptr, nullptr}, as_stream(stream));
}

// θ̄ = (∂f/∂θ at the prepared point)ᵀ ȳ as a flat vector in the layout of the packed weight block (psi_weights_floats() floats, only
// the LayerWeights part is written); also returns Jᵀȳ (dev_jty, may be NULL).  tab_*: the table of psi_gnn_b200/weights.py.
extern "C" int psi_param_grad(psi_graph_t* g, int kind, const float* dev_hstar, const float* dev_ybar, const int32_t* dev_tab_dst,
                              const int32_t* dev_tab_y, const int32_t* dev_tab_x, int n_tab, float* dev_out, float* dev_jty, void* stream) {
    if (check_kind(g, kind)) return -1;
    if (kind != PSI_KIND_DIRICHLET && kind != PSI_KIND_MIXED) PSI_FAIL("psi_param_grad: implemented for the PSI-GNN layers");
    if (!g->vjp_ready || g->vjp_kind != kind) PSI_FAIL("psi_param_grad: call psi_vjp_prepare (at the same H*) first");
    if (n_tab < 1 || n_tab > PG_NODES * PG_MAX_PER_THREAD) PSI_FAIL("psi_param_grad: table size out of range");
    if (dev_out == nullptr || dev_tab_dst == nullptr || dev_tab_y == nullptr || dev_tab_x == nullptr) PSI_FAIL("psi_param_grad: null pointer");
    cudaStream_t st = as_stream(stream);
    PSI_CK(cudaMemsetAsync(dev_out, 0, sizeof(LayerWeights), st));
    if (g->N == 0) return 0;
    if (dev_hstar == nullptr || dev_ybar == nullptr) PSI_FAIL("psi_param_grad: null pointer");
    const float* hs = (g->part != nullptr && g->p_hstar != nullptr) ? g->p_hstar : dev_hstar;
    float *acc = nullptr, *partial = nullptr, *jty = dev_jty;
    PSI_CK(psi_malloc_async((void**)&acc, (size_t)g->N * 30 * sizeof(float), st));
    if (jty == nullptr) PSI_CK(psi_malloc_async((void**)&jty, (size_t)g->N * PSI_D * sizeof(float), st));
    const SolverEpi noE{nullptr, nullptr, nullptr, nullptr};
    int rc = launch_vjp<false>(g, kind, dev_ybar, nullptr, jty, noE, st, acc);
    const int num_batches = (int)((g->dev.n_compute + PG_NODES - 1) / PG_NODES);
    const int grid = std::max(1, std::min(num_batches, PSI_NUM_SMS_B200 * 2));
    if (!rc && psi_malloc_async((void**)&partial, (size_t)grid * n_tab * sizeof(float), st) != cudaSuccess) { g_psi_err = "psi_param_grad: out of device memory"; rc = -1; }
    if (!rc) {
        const size_t smem = (size_t)PG_NODES * PG_PITCH * sizeof(float);
        static bool attr_done = false;
        if (!attr_done) {
            cudaFuncSetAttribute(k_pgrad<KIND_DIRICHLET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k_pgrad<KIND_MIXED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr_done = true;
        }
        if (kind == PSI_KIND_DIRICHLET)
            k_pgrad<KIND_DIRICHLET><<<grid, PG_NODES, smem, st>>>(g->dev, g->vjp, hs, dev_ybar, acc, dev_tab_y, dev_tab_x, n_tab, partial, num_batches);
        else
            k_pgrad<KIND_MIXED><<<grid, PG_NODES, smem, st>>>(g->dev, g->vjp, hs, dev_ybar, acc, dev_tab_y, dev_tab_x, n_tab, partial, num_batches);
        k_pgrad_reduce<<<(n_tab + 127) / 128, 128, 0, st>>>(partial, grid, n_tab, dev_tab_dst, dev_out);
        if (cudaGetLastError() != cudaSuccess) { g_psi_err = "psi_param_grad: kernel launch failed"; rc = -1; }
    }
    psi_free_async(acc, st);
    psi_free_async(partial, st);
    if (dev_jty == nullptr) psi_free_async(jty, st);
    return rc;
}

// θ̄' = d/dε θ̄(H* + ε ḣ; ȳ) at ε = 0: the tangent of psi_param_grad along a direction of the frozen point (pgrad.cuh, second half).
// With ȳ = v (the Hutchinson probe) and ḣ = Jᵀv this is ½ ∇θ ‖Jᵀv‖² — the double backward of jac_loss_estimate
// (dirichlet/psignn/model.py:207, :416-435).  Same table, same output layout, deterministic.
extern "C" int psi_param_grad_tangent(psi_graph_t* g, int kind, const float* dev_hstar, const float* dev_ybar, const float* dev_hdot,
                                      const int32_t* dev_tab_dst, const int32_t* dev_tab_y, const int32_t* dev_tab_x, int n_tab,
                                      float* dev_out, void* stream) {
    if (check_kind(g, kind)) return -1;
    if (kind != PSI_KIND_DIRICHLET && kind != PSI_KIND_MIXED) PSI_FAIL("psi_param_grad_tangent: implemented for the PSI-GNN layers");
    if (!g->vjp_ready || g->vjp_kind != kind) PSI_FAIL("psi_param_grad_tangent: call psi_vjp_prepare (at the same H*) first");
    if (g->part != nullptr) PSI_FAIL("psi_param_grad_tangent: not available on a mesh partition (training shards whole graphs)");
    if (n_tab < 1 || n_tab > PG_NODES * PG_MAX_PER_THREAD) PSI_FAIL("psi_param_grad_tangent: table size out of range");
    if (dev_out == nullptr || dev_tab_dst == nullptr || dev_tab_y == nullptr || dev_tab_x == nullptr) PSI_FAIL("psi_param_grad_tangent: null pointer");
    cudaStream_t st = as_stream(stream);
    PSI_CK(cudaMemsetAsync(dev_out, 0, sizeof(LayerWeights), st));
    if (g->N == 0) return 0;
    if (dev_hstar == nullptr || dev_ybar == nullptr || dev_hdot == nullptr) PSI_FAIL("psi_param_grad_tangent: null pointer");
    const int num_batches = (int)((g->dev.n_compute + PG_NODES - 1) / PG_NODES);
    const int grid = std::max(1, std::min(num_batches, PSI_NUM_SMS_B200));
    const size_t n30 = (size_t)g->N * 30, nsb = (size_t)2 * g->N * PSI_QPITCH, nrow = (size_t)g->N * PSI_D;
    float* buf = nullptr;     // acc | acc' | S̄' | scratch row output of phase B | per-CTA partials
    PSI_CK(psi_malloc_async((void**)&buf, (2 * n30 + nsb + nrow + (size_t)grid * n_tab) * sizeof(float), st));
    float *acc = buf, *acc_t = acc + n30, *sb_t = acc_t + n30, *scratch = sb_t + nsb, *partial = scratch + nrow;
    const SolverEpi noE{nullptr, nullptr, nullptr, nullptr};
    int rc = launch_vjp<false>(g, kind, dev_ybar, nullptr, scratch, noE, st, acc);            // acc of the primal cotangents
    if (!rc) {
        const size_t smem = (size_t)2 * PG_NODES * PG_PITCH * sizeof(float);
        static bool attr_done = false;
        if (!attr_done) {
            cudaFuncSetAttribute(k_pgrad_tan<KIND_DIRICHLET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k_pgrad_tan<KIND_MIXED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr_done = true;
        }
        VjpCacheDev Ct = g->vjp;                   // same masks and statistics; the gathered array is S̄' instead of S̄
        Ct.Sb = sb_t;
        const unsigned ngrid = node_grid(g->dev.n_compute);
        if (kind == PSI_KIND_DIRICHLET) {
            k_pgrad_tan<KIND_DIRICHLET><<<grid, PG_NODES, smem, st>>>(g->dev, g->vjp, dev_hstar, dev_hdot, dev_ybar, acc, nullptr, dev_tab_y, dev_tab_x, n_tab, partial, num_batches, sb_t, 0);
            k_vjp_phase_b<KIND_DIRICHLET, false><<<ngrid, PSI_NODE_BLOCK, 0, st>>>(g->dev, Ct, dev_ybar, nullptr, scratch, noE, acc_t);
            k_pgrad_tan<KIND_DIRICHLET><<<grid, PG_NODES, smem, st>>>(g->dev, g->vjp, dev_hstar, dev_hdot, dev_ybar, acc, acc_t, dev_tab_y, dev_tab_x, n_tab, partial, num_batches, nullptr, 1);
        } else {
            k_pgrad_tan<KIND_MIXED><<<grid, PG_NODES, smem, st>>>(g->dev, g->vjp, dev_hstar, dev_hdot, dev_ybar, acc, nullptr, dev_tab_y, dev_tab_x, n_tab, partial, num_batches, sb_t, 0);
            k_vjp_phase_b<KIND_MIXED, false><<<ngrid, PSI_NODE_BLOCK, 0, st>>>(g->dev, Ct, dev_ybar, nullptr, scratch, noE, acc_t);
            k_pgrad_tan<KIND_MIXED><<<grid, PG_NODES, smem, st>>>(g->dev, g->vjp, dev_hstar, dev_hdot, dev_ybar, acc, acc_t, dev_tab_y, dev_tab_x, n_tab, partial, num_batches, nullptr, 1);
        }
        k_pgrad_reduce<<<(n_tab + 127) / 128, 128, 0, st>>>(partial, grid, n_tab, dev_tab_dst, dev_out);
        if (cudaGetLastError() != cudaSuccess) { g_psi_err = "psi_param_grad_tangent: kernel launch failed"; rc = -1; }
    }
    psi_free_async(buf, st);
    return rc;
}

// Backward of ONE unrolled layer of the DSS / DSGPS baselines at its own input h (baseline_bwd.cuh): h̄ = Jᵀȳ and the parameter gradient
// θ̄ in packed-block layout.  The layer's weight block must be resident (psi_weights_upload).  Deterministic, no atomics.
template <int KIND>
static int launch_baseline_backward(psi_graph* g, const float* h, const float* ybar, const int32_t* tab_dst, const int32_t* tab_y,
                                    const int32_t* tab_x, int n_tab, float* hbar, float* out, cudaStream_t st) {
    const int64_t N = g->N;
    float *Dloc = nullptr, *Sb = nullptr, *acc = nullptr, *partial = nullptr;
    const int num_batches = (int)((g->dev.n_compute + PG_NODES - 1) / PG_NODES);
    const int grid = std::max(1, std::min(num_batches, PSI_NUM_SMS_B200 * 2));
    const size_t smem = (size_t)PG_NODES * PG_PITCH * sizeof(float);
    int rc = 0;
    if (psi_malloc_async((void**)&Dloc, (size_t)N * PSI_D * sizeof(float), st) != cudaSuccess ||
        psi_malloc_async((void**)&Sb, (size_t)2 * N * PSI_QPITCH * sizeof(float), st) != cudaSuccess ||
        psi_malloc_async((void**)&acc, (size_t)N * 30 * sizeof(float), st) != cudaSuccess ||
        psi_malloc_async((void**)&partial, (size_t)grid * n_tab * sizeof(float), st) != cudaSuccess) {
        g_psi_err = "psi_layer_backward: out of device memory";
        rc = -1;
    }
    if (!rc) {
        static bool attr_done = false;
        if (!attr_done) {
            cudaFuncSetAttribute(k_bl_pass<KIND_DSS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k_bl_pass<KIND_DSGPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k_bl_pass<KIND_DSGPS_MIXED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr_done = true;
        }
        // PSI_BL_DEBUG=1: synchronise behind every kernel and name the one that failed
        static const bool dbg = getenv("PSI_BL_DEBUG") != nullptr;
        auto stage = [&](const char* what) {
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess && dbg) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess && !rc) { g_psi_err = std::string("psi_layer_backward: ") + what + " -> " + cudaGetErrorString(e); rc = -1; }
        };
        k_bl_pass<KIND><<<grid, PG_NODES, smem, st>>>(g->dev, h, ybar, nullptr, Dloc, Sb, tab_y, tab_x, n_tab, partial, num_batches, 0);
        stage("k_bl_pass(0)");
        k_bl_gather<KIND><<<node_grid(g->dev.n_compute), PSI_NODE_BLOCK, 0, st>>>(g->dev, h, Dloc, Sb, hbar, acc);
        stage("k_bl_gather");
        k_bl_pass<KIND><<<grid, PG_NODES, smem, st>>>(g->dev, h, ybar, acc, Dloc, Sb, tab_y, tab_x, n_tab, partial, num_batches, 1);
        stage("k_bl_pass(1)");
        k_pgrad_reduce<<<(n_tab + 127) / 128, 128, 0, st>>>(partial, grid, n_tab, tab_dst, out);
        stage("k_pgrad_reduce");
    }
    psi_free_async(Dloc, st);
    psi_free_async(Sb, st);
    psi_free_async(acc, st);
    psi_free_async(partial, st);
    return rc;
}

extern "C" int psi_layer_backward(psi_graph_t* g, int kind, const float* dev_h, const float* dev_ybar, const int32_t* dev_tab_dst,
                                  const int32_t* dev_tab_y, const int32_t* dev_tab_x, int n_tab, float* dev_hbar, float* dev_out, void* stream) {
    if (check_kind(g, kind)) return -1;
    if (kind != PSI_KIND_DSS && kind != PSI_KIND_DSGPS && kind != PSI_KIND_DSGPS_MIXED)
        PSI_FAIL("psi_layer_backward: for the unrolled baseline layers (the PSI-GNN layers use psi_vjp_apply / psi_param_grad)");
    if (g->part != nullptr) PSI_FAIL("psi_layer_backward: not available on a mesh partition");
    if (n_tab < 1 || n_tab > PG_NODES * BL_MAX_PER_THREAD) PSI_FAIL("psi_layer_backward: table size out of range");
    if (dev_out == nullptr || dev_tab_dst == nullptr || dev_tab_y == nullptr || dev_tab_x == nullptr) PSI_FAIL("psi_layer_backward: null pointer");
    cudaStream_t st = as_stream(stream);
    PSI_CK(cudaMemsetAsync(dev_out, 0, sizeof(LayerWeights), st));
    if (g->N == 0) return 0;
    if (dev_h == nullptr || dev_ybar == nullptr || dev_hbar == nullptr) PSI_FAIL("psi_layer_backward: null pointer");
    if (kind == PSI_KIND_DSS) return launch_baseline_backward<KIND_DSS>(g, dev_h, dev_ybar, dev_tab_dst, dev_tab_y, dev_tab_x, n_tab, dev_hbar, dev_out, st);
    if (kind == PSI_KIND_DSGPS) return launch_baseline_backward<KIND_DSGPS>(g, dev_h, dev_ybar, dev_tab_dst, dev_tab_y, dev_tab_x, n_tab, dev_hbar, dev_out, st);
    return launch_baseline_backward<KIND_DSGPS_MIXED>(g, dev_h, dev_ybar, dev_tab_dst, dev_tab_y, dev_tab_x, n_tab, dev_hbar, dev_out, st);
}

// record layout of pgrad.cuh for the table builder: {PG_ONE, PG_DEG, PG_C, PG_CN, PG_YB, PG_RHAT, PG_MB, PG_HID, PG_TB, PG_SB, PG_EDGE,
// PG_ACC, PG_MBN, PG_HIDN, PG_TBN, PG_REC}
extern "C" int psi_pgrad_layout(int32_t out[16]) {
    const int32_t v[16] = {PG_ONE, PG_DEG, PG_C, PG_CN, PG_YB, PG_RHAT, PG_MB, PG_HID, PG_TB, PG_SB, PG_EDGE, PG_ACC, PG_MBN, PG_HIDN, PG_TBN, PG_REC};
    for (int i = 0; i < 16; ++i) out[i] = v[i];
    return 0;
}

extern "C" int psi_residual(const psi_graph_t* g, const float* dev_u, const float* dev_y, float* dev_r, float* dev_mean_sq, void* stream) {
    if (g == nullptr) PSI_FAIL("psi_residual: null graph");
    if (g->p_recs_Ar == nullptr) PSI_FAIL("psi_residual: graph was created without a_ij");
    cudaStream_t st = as_stream(stream);
    if (g->N == 0) {
        if (dev_mean_sq) {   // mean over an empty tensor is NaN in torch
            const float nanv = std::numeric_limits<float>::quiet_NaN();
            PSI_CK(cudaMemcpyAsync(dev_mean_sq, &nanv, sizeof(float), cudaMemcpyHostToDevice, st));
            PSI_CK(cudaStreamSynchronize(st));
        }
        return 0;
    }
    if (dev_u == nullptr || dev_y == nullptr) PSI_FAIL("psi_residual: null pointer");
    const unsigned grid = node_grid(g->N);
    k_residual<<<grid, PSI_NODE_BLOCK, 0, st>>>(g->dev, dev_u, dev_y, dev_r, g->p_scratch);
    PSI_CK_LAUNCH();
    if (dev_mean_sq != nullptr) {
        k_reduce_partials<<<1, 256, 0, st>>>((int)grid, g->p_scratch, 1.0f / (float)g->N, dev_mean_sq);
        PSI_CK_LAUNCH();
    }
    return 0;
}

extern "C" int psi_spmv_t(const psi_graph_t* g, const float* dev_v, float* dev_out, void* stream) {
    if (g == nullptr) PSI_FAIL("psi_spmv_t: null graph");
    if (g->p_recs_Ac == nullptr) PSI_FAIL("psi_spmv_t: graph was created without a_ij");
    if (g->N == 0) return 0;
    if (dev_v == nullptr || dev_out == nullptr) PSI_FAIL("psi_spmv_t: null pointer");
    k_spmv_t<<<node_grid(g->N), PSI_NODE_BLOCK, 0, as_stream(stream)>>>(g->dev, dev_v, dev_out);
    PSI_CK_LAUNCH();
    return 0;
}

extern "C" int psi_flux(const psi_graph_t* g, const float* dev_v, float* dev_out, int transpose, void* stream) {
    if (g == nullptr) PSI_FAIL("psi_flux: null graph");
    if (g->p_recs_Ar == nullptr || g->p_recs_Ac == nullptr) PSI_FAIL("psi_flux: graph was created without a_ij");
    if (g->N == 0) return 0;
    if (dev_v == nullptr || dev_out == nullptr || dev_v == dev_out) PSI_FAIL("psi_flux: null or aliased pointer");
    if (transpose) k_flux<true><<<node_grid(g->N), PSI_NODE_BLOCK, 0, as_stream(stream)>>>(g->dev, dev_v, dev_out);
    else k_flux<false><<<node_grid(g->N), PSI_NODE_BLOCK, 0, as_stream(stream)>>>(g->dev, dev_v, dev_out);
    PSI_CK_LAUNCH();
    return 0;
}

extern "C" int psi_encode(int64_t num_nodes, const float* dev_x, float* dev_h, void* stream) {
    if (num_nodes < 0) PSI_FAIL("psi_encode: negative size");
    if (num_nodes == 0) return 0;
    if (dev_x == nullptr || dev_h == nullptr) PSI_FAIL("psi_encode: null pointer");
    k_encode<<<node_grid(num_nodes), PSI_NODE_BLOCK, 0, as_stream(stream)>>>((int)num_nodes, dev_x, dev_h);
    PSI_CK_LAUNCH();
    return 0;
}

extern "C" int psi_decode(int64_t num_nodes, const float* dev_h, float* dev_u, void* stream) {
    if (num_nodes < 0) PSI_FAIL("psi_decode: negative size");
    if (num_nodes == 0) return 0;
    if (dev_h == nullptr || dev_u == nullptr) PSI_FAIL("psi_decode: null pointer");
    k_decode<<<node_grid(num_nodes), PSI_NODE_BLOCK, 0, as_stream(stream)>>>((int)num_nodes, dev_h, dev_u);
    PSI_CK_LAUNCH();
    return 0;
}

// ================================================================================================
// solver workspace
// ================================================================================================
#define PROF_CLASSES 3   // 0 operator (layer / VJP pair), 1 k_qn_dots_tma, 2 k_qn_axpy_tma

struct psi_solver {
    int64_t numel = 0, stride = 0;        // stride = numel rounded up to whole QN_CHUNKs (tail kept at zero)
    int cap = 0;                          // largest threshold the workspace can hold
    int num_chunks = 0, axpy_ctas = 0, norm_cap = 0;
    int dots_chunks = 0, tma_ctas = 0;    // 2048-element chunks of pass 1; persistent CTAs (one per SM) of the TMA kernels
    float *x = nullptr, *g = nullptr, *dg = nullptr, *dx = nullptr, *best = nullptr, *fx = nullptr;
    float *partial = nullptr, *coef = nullptr, *norm_part = nullptr;
    QnCtrl* ctrl = nullptr;               // device
    QnCtrl* h_ctrl = nullptr;             // pinned host mirror
    int* h_done = nullptr;                // mapped pinned word the device sets when a solve stops (read by the host without a sync)
    int* d_done_host = nullptr;           // its device-side address
    cudaEvent_t ahead[4] = {nullptr, nullptr, nullptr, nullptr};   // run-ahead window of the fused loops
    double *rel_trace = nullptr, *abs_trace = nullptr;
    QnHistory hist{};
    int slabs_alloc = 0;
    int64_t bytes = 0;
    // Anderson window (allocated on first use)
    float* and_X = nullptr; float* and_F = nullptr; float* and_small = nullptr; int and_m = 0;
    int and_cur_m = 2, and_e = 0; double and_lam = 1e-4, and_beta = 1.0;   // state of the Anderson / Picard step machines
    // state of the step API
    int threshold = 0; double eps = 0.0; int n = 0; int launches = 0; int f_evals = 0; bool active = false;
    float* xtrace = nullptr; int norm_blocks = 0;
    // active extent of the current solve: all of numel, or the owned rows of a mesh partition (ghost rows follow in x)
    int64_t act_numel = 0, last_act = -1; int act_chunks = 0, act_dchunks = 0;
    double* dbuf = nullptr;               // [3·cap + 8] fp64 sums all-reduced over the ranks of a partitioned solve
    psi_comm* comm = nullptr;             // communicator of the current solve (nullptr: single rank)
    Partition* part = nullptr;            // partition of the current solve (peer-mapped mailboxes when part->p2p)
    // optional per-kernel-class timing with CUDA events on the launching stream (psi_solver_profile)
    int profile = 0;
    std::vector<cudaEvent_t> ev;          // [((step * PROF_CLASSES) + cls) * 2 + {begin,end}]
    std::vector<double> ev_bytes;         // algorithmic bytes of the launch(es) bracketed by the pair
    double prof_ms[PROF_CLASSES] = {0, 0, 0}, prof_bytes[PROF_CLASSES] = {0, 0, 0};
    int64_t prof_launches[PROF_CLASSES] = {0, 0, 0};
    double op_bytes = 0.0;                // algorithmic bytes of one operator evaluation of the current solve
};

static void prof_begin(psi_solver* s, int step, int cls, double bytes, cudaStream_t st) {
    if (!s->profile) return;
    const size_t i = ((size_t)step * PROF_CLASSES + cls) * 2;
    if (i + 1 >= s->ev.size()) return;
    s->ev_bytes[i / 2] = bytes;
    cudaEventRecord(s->ev[i], st);
}
static void prof_end(psi_solver* s, int step, int cls, cudaStream_t st) {
    if (!s->profile) return;
    const size_t i = ((size_t)step * PROF_CLASSES + cls) * 2;
    if (i + 1 >= s->ev.size()) return;
    cudaEventRecord(s->ev[i + 1], st);
}
// fold the event pairs of steps 0..last_step into the running totals (stream must be synchronised)
static void prof_collect(psi_solver* s, int last_step) {
    if (!s->profile) return;
    for (int step = 0; step <= last_step; ++step)
        for (int cls = 0; cls < PROF_CLASSES; ++cls) {
            const size_t i = ((size_t)step * PROF_CLASSES + cls) * 2;
            if (i + 1 >= s->ev.size() || s->ev_bytes[i / 2] < 0.0) continue;
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, s->ev[i], s->ev[i + 1]) == cudaSuccess) {
                s->prof_ms[cls] += ms; s->prof_bytes[cls] += s->ev_bytes[i / 2]; s->prof_launches[cls] += 1;
            }
        }
    (void)cudaGetLastError();
    std::fill(s->ev_bytes.begin(), s->ev_bytes.end(), -1.0);
}

static int solver_alloc(psi_solver* s, void** p, size_t bytes) {
    PSI_CK(cudaMalloc(p, bytes));
    s->bytes += (int64_t)bytes;
    return 0;
}

extern "C" int psi_solver_create(psi_solver_t** out, int64_t numel, int max_threshold) {
    if (out == nullptr) PSI_FAIL("psi_solver_create: null out");
    *out = nullptr;
    if (numel < 0 || max_threshold < 1) PSI_FAIL("psi_solver_create: bad size");
    psi_solver* s = new psi_solver();
    s->numel = numel;
    s->stride = round_up64(numel > 0 ? numel : 1, 4096);
    s->cap = max_threshold;
    s->num_chunks = (int)(s->stride / QN_CHUNK);
    s->axpy_ctas = std::min(s->num_chunks, QN_AXPY_MAX_CTAS);
    s->dots_chunks = (int)(s->stride / DOTS_CH);
    {
        int dev = 0, sms = PSI_NUM_SMS_B200;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = PSI_NUM_SMS_B200;
        s->tma_ctas = std::min(sms, QN_AXPY_MAX_CTAS);
        static bool attr_done = false;
        if (!attr_done) {
            cudaFuncSetAttribute(k_qn_dots_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dots_tma_smem());
            cudaFuncSetAttribute(k_qn_axpy_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)axpy_tma_smem(4096));
            attr_done = true;
        }
    }
    s->norm_cap = std::max((int)node_grid(numel / PSI_D + 1), s->num_chunks) + 1;
    const size_t vb = s->stride * sizeof(float);
    int rc = 0;
    rc |= solver_alloc(s, (void**)&s->x, vb);
    rc |= solver_alloc(s, (void**)&s->g, vb);
    rc |= solver_alloc(s, (void**)&s->dg, vb);
    rc |= solver_alloc(s, (void**)&s->dx, vb);
    rc |= solver_alloc(s, (void**)&s->best, vb);
    rc |= solver_alloc(s, (void**)&s->fx, vb);
    rc |= solver_alloc(s, (void**)&s->partial, (size_t)(3 * s->cap + 2) * s->dots_chunks * sizeof(float));
    rc |= solver_alloc(s, (void**)&s->coef, (size_t)3 * s->cap * sizeof(float));
    rc |= solver_alloc(s, (void**)&s->norm_part, (size_t)2 * s->norm_cap * sizeof(float));
    rc |= solver_alloc(s, (void**)&s->ctrl, sizeof(QnCtrl));
    rc |= solver_alloc(s, (void**)&s->dbuf, (size_t)(2 * (3 * s->cap + 8) + 16) * sizeof(double));
    rc |= solver_alloc(s, (void**)&s->rel_trace, (size_t)(s->cap + 2) * sizeof(double));
    rc |= solver_alloc(s, (void**)&s->abs_trace, (size_t)(s->cap + 2) * sizeof(double));
    if (!rc && cudaMallocHost((void**)&s->h_ctrl, sizeof(QnCtrl)) != cudaSuccess) { g_psi_err = "psi_solver_create: pinned allocation failed"; rc = -1; }
    if (!rc && (cudaHostAlloc((void**)&s->h_done, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
                cudaHostGetDevicePointer((void**)&s->d_done_host, s->h_done, 0) != cudaSuccess)) { g_psi_err = "psi_solver_create: mapped allocation failed"; rc = -1; }
    for (int i = 0; i < 4 && !rc; ++i)
        if (cudaEventCreateWithFlags(&s->ahead[i], cudaEventDisableTiming) != cudaSuccess) { g_psi_err = "psi_solver_create: event creation failed"; rc = -1; }
    if (rc) { psi_solver_destroy(s); return -1; }
    float* vecs[] = {s->x, s->g, s->dg, s->dx, s->best, s->fx};
    for (float* v : vecs) cudaMemset(v, 0, vb);
    // history slabs: at most QN_MAX_SLABS, each at least ~64 MB so that small problems use a single slab
    int sv = (s->cap + QN_MAX_SLABS - 1) / QN_MAX_SLABS;
    const int64_t min_vecs = std::max<int64_t>(1, (64ll << 20) / (int64_t)vb);
    if (sv < min_vecs) sv = (int)std::min<int64_t>(min_vecs, s->cap);
    s->hist.slab_vecs = sv;
    s->hist.stride = s->stride;
    for (int i = 0; i < QN_MAX_SLABS; ++i) { s->hist.U[i] = nullptr; s->hist.V[i] = nullptr; }
    *out = s;
    return 0;
}

extern "C" int psi_solver_destroy(psi_solver_t* s) {
    if (s == nullptr) return 0;
    void* ps[] = {s->x, s->g, s->dg, s->dx, s->best, s->fx, s->partial, s->coef, s->norm_part, s->ctrl,
                  s->rel_trace, s->abs_trace, s->and_X, s->and_F, s->and_small, s->dbuf};
    for (void* p : ps)
        if (p) cudaFree(p);
    for (int i = 0; i < QN_MAX_SLABS; ++i) {
        if (s->hist.U[i]) cudaFreeAsync(s->hist.U[i], nullptr);
        if (s->hist.V[i]) cudaFreeAsync(s->hist.V[i], nullptr);
    }
    if (s->h_ctrl) cudaFreeHost(s->h_ctrl);
    if (s->h_done) cudaFreeHost(s->h_done);
    for (int i = 0; i < 4; ++i) if (s->ahead[i]) cudaEventDestroy(s->ahead[i]);
    for (cudaEvent_t e : s->ev) cudaEventDestroy(e);
    delete s;
    return 0;
}

extern "C" int64_t psi_solver_bytes(const psi_solver_t* s) { return s ? s->bytes : 0; }
extern "C" int64_t psi_solver_stride(const psi_solver_t* s) { return s ? s->stride : 0; }

extern "C" int psi_solver_profile(psi_solver_t* s, int enable) {
    if (s == nullptr) PSI_FAIL("psi_solver_profile: null solver");
    if (enable && s->ev.empty()) {
        const size_t n = (size_t)(s->cap + 2) * PROF_CLASSES * 2;
        s->ev.resize(n);
        for (size_t i = 0; i < n; ++i) PSI_CK(cudaEventCreate(&s->ev[i]));
        s->ev_bytes.assign(n / 2, -1.0);
    }
    s->profile = enable ? 1 : 0;
    if (enable)
        for (int c = 0; c < PROF_CLASSES; ++c) { s->prof_ms[c] = 0; s->prof_bytes[c] = 0; s->prof_launches[c] = 0; }
    return 0;
}

extern "C" int psi_solver_profile_read(const psi_solver_t* s, double out[9]) {
    if (s == nullptr || out == nullptr) PSI_FAIL("psi_solver_profile_read: null argument");
    for (int c = 0; c < PROF_CLASSES; ++c) { out[3 * c] = (double)s->prof_launches[c]; out[3 * c + 1] = s->prof_ms[c]; out[3 * c + 2] = s->prof_bytes[c]; }
    return 0;
}

// make sure history vector index k (0-based) has storage.  Slabs come from the stream-ordered pool and are zeroed on the solve's
// stream: growing the history inside a running loop costs no device synchronisation (a plain cudaMalloc would drain the queue)
static int hist_ensure(psi_solver* s, int k, cudaStream_t st) {
    const int slab = k / s->hist.slab_vecs;
    if (slab >= QN_MAX_SLABS) PSI_FAIL("solver history exhausted");
    while (s->slabs_alloc <= slab) {
        const int first = s->slabs_alloc * s->hist.slab_vecs;
        const int vecs = std::min(s->hist.slab_vecs, s->cap - first);
        if (vecs <= 0) PSI_FAIL("solver history exhausted");
        const size_t b = (size_t)vecs * s->stride * sizeof(float);
        PSI_CK(psi_malloc_async((void**)&s->hist.U[s->slabs_alloc], b, st));
        PSI_CK(psi_malloc_async((void**)&s->hist.V[s->slabs_alloc], b, st));
        s->bytes += 2 * (int64_t)b;
        // the padding beyond the active extent is read by the streaming kernels (times zero) and never written: keep it finite
        PSI_CK(cudaMemsetAsync(s->hist.U[s->slabs_alloc], 0, b, st));
        PSI_CK(cudaMemsetAsync(s->hist.V[s->slabs_alloc], 0, b, st));
        ++s->slabs_alloc;
    }
    return 0;
}

// ---- shared pieces of the Broyden loop ------------------------------------------------------------------
static int qn_begin(psi_solver* s, const float* x0, int threshold, double eps, float* xtrace, cudaStream_t st, int64_t act_numel = -1,
                    psi_comm* comm = nullptr) {
    if (s == nullptr) PSI_FAIL("null solver handle");
    if (threshold < 0 || threshold > s->cap) PSI_FAIL("threshold exceeds the solver workspace (psi_solver_create max_threshold)");
    if (s->numel > 0 && x0 == nullptr) PSI_FAIL("null x0");
    s->threshold = threshold; s->eps = eps; s->n = 0; s->launches = 0; s->f_evals = 0; s->xtrace = xtrace; s->active = true;
    s->comm = (comm != nullptr && comm->world > 1) ? comm : nullptr;
    s->part = nullptr;
    s->act_numel = (act_numel < 0 || act_numel > s->numel) ? s->numel : act_numel;
    s->act_chunks = (int)((s->act_numel + QN_CHUNK - 1) / QN_CHUNK);
    s->act_dchunks = (int)((s->act_numel + DOTS_CH - 1) / DOTS_CH);
    if (s->last_act != s->act_numel) {
        // a different extent than the previous solve on this workspace: re-establish the zero padding the kernels rely on
        const size_t vb = s->stride * sizeof(float);
        PSI_CK(cudaMemsetAsync(s->g, 0, vb, st)); PSI_CK(cudaMemsetAsync(s->dg, 0, vb, st)); PSI_CK(cudaMemsetAsync(s->dx, 0, vb, st));
        PSI_CK(cudaMemsetAsync(s->x, 0, vb, st)); PSI_CK(cudaMemsetAsync(s->best, 0, vb, st));
        for (int i = 0; i < s->slabs_alloc; ++i) {
            const int first = i * s->hist.slab_vecs;
            const size_t b = (size_t)std::min(s->hist.slab_vecs, s->cap - first) * vb;
            PSI_CK(cudaMemsetAsync(s->hist.U[i], 0, b, st)); PSI_CK(cudaMemsetAsync(s->hist.V[i], 0, b, st));
        }
        s->last_act = s->act_numel;
    }
    *s->h_done = 0;
    k_qn_ctrl_init<<<1, 1, 0, st>>>(s->ctrl, s->d_done_host);
    PSI_CK_LAUNCH();
    if (s->numel > 0) {
        PSI_CK(cudaMemcpyAsync(s->x, x0, s->numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
        // lowest_xest starts as x0 (solver.py:150)
        PSI_CK(cudaMemcpyAsync(s->best, x0, s->numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
        if (xtrace != nullptr) PSI_CK(cudaMemcpyAsync(xtrace, s->x, s->stride * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    s->launches += 1;
    return 0;
}

// after g_0 is in s->g: update_0 = g_0, x_1 = x_0 + update_0
static int qn_first(psi_solver* s, cudaStream_t st) {
    if (s->threshold == 0) return 0;
    k_qn_first<<<s->axpy_ctas, QN_THREADS, 0, st>>>(s->dx, s->g, s->x, s->xtrace ? s->xtrace + s->stride : nullptr, s->act_chunks);
    PSI_CK_LAUNCH();
    s->launches += 1;
    return 0;
}

// bookkeeping of step n (1-based) after the operator epilogue produced g_n, δg and the norm partials
#ifdef PSI_DOTS_SELFCHECK
static inline float __int_as_float_host(int v) { float f; memcpy(&f, &v, 4); return f; }
__global__ void k_dots_selfcheck(const float* a, const float* b, int count, int chunks, int step, int kr, int* dbg, const int* done) {
    if (*done) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        if (__float_as_int(a[i]) != __float_as_int(b[i])) {
            if (atomicAdd(&dbg[0], 1) == 0) {
                dbg[1] = step; dbg[2] = i / chunks; dbg[3] = i % chunks; dbg[4] = kr; dbg[5] = __float_as_int(a[i]); dbg[6] = __float_as_int(b[i]);
                dbg[7] = (count / chunks - 2) / 3;
            }
        }
    }
}
#endif

static int qn_update(psi_solver* s, int n, int norm_blocks, cudaStream_t st) {
    const int nhist = n - 1;
    if (hist_ensure(s, n - 1, st)) return -1;
    const double vec = (double)s->act_numel * 4.0;
    // pass 1 (also at nhist = 0: it carries ⟨δx,δg⟩ and ⟨δx,g⟩, from which s and p of this step follow)
    prof_begin(s, n, 1, (2.0 * nhist + 3.0) * vec, st);
    // history vectors per work item: 32, or — for small problems — the finest of {16, 8} whose items still fit ONE wave of CTAs
    // (more SMs busy; a second wave of single-batch items is avoided: repeated long solves at 7 … 15 chunks were not bitwise
    // repeatable in that regime, see NOTES.md)
    int kr = DOTS_KR;
#if defined(PSI_KR_FINE)
    while (kr > 8 && (int64_t)s->act_dchunks * ((nhist + kr - 1) / kr) < 2 * (int64_t)s->tma_ctas) kr >>= 1;
#elif !defined(PSI_FIXED_KR)
    for (int cand = 16; cand >= 8; cand >>= 1)
        if ((int64_t)s->act_dchunks * ((nhist + cand - 1) / cand) <= (int64_t)s->tma_ctas) kr = cand;
#endif
#ifdef PSI_FORCE_KR
    kr = PSI_FORCE_KR;
#endif
    k_qn_dots_tma<<<s->tma_ctas, TMA_THREADS, dots_tma_smem(), st>>>(s->hist, nhist, s->dx, s->dg, s->g, s->partial, s->act_dchunks,
                                                                     &s->ctrl->done, kr);
    prof_end(s, n, 1, st);
    PSI_CK_LAUNCH();
#ifdef PSI_DOTS_SELFCHECK
    {   // diagnosis: the same pass with 32-vector items into a second buffer; the partial sums must agree bit for bit
        static float* p2 = nullptr;
        static int* dbg = nullptr;
        static size_t p2_floats = 0;
        const size_t need = (size_t)(3 * s->cap + 2) * (size_t)s->num_chunks;
        if (p2 == nullptr || p2_floats < need) { cudaMalloc(&p2, need * sizeof(float)); p2_floats = need; }
        if (dbg == nullptr) { cudaMallocManaged(&dbg, 64 * sizeof(int)); memset(dbg, 0, 64 * sizeof(int)); }
        k_qn_dots_tma<<<s->tma_ctas, TMA_THREADS, dots_tma_smem(), st>>>(s->hist, nhist, s->dx, s->dg, s->g, p2, s->act_dchunks, &s->ctrl->done, DOTS_KR);
        k_dots_selfcheck<<<64, 256, 0, st>>>(s->partial, p2, (3 * nhist + 2) * s->act_dchunks, s->act_dchunks, n, kr, dbg, &s->ctrl->done);
        if (n == s->threshold || n % 50 == 0) {
            cudaStreamSynchronize(st);
            if (dbg[0] > 0 && dbg[8] == 0) {
                dbg[8] = 1;
                fprintf(stderr, "[selfcheck] %d mismatching partials so far; first at step %d: row %d (k %d q %d) chunk %d kr %d: %.9g vs %.9g (nhist %d chunks %d)\n",
                        dbg[0], dbg[1], dbg[2], dbg[2] / 3, dbg[2] % 3, dbg[3], dbg[4], __int_as_float_host(dbg[5]), __int_as_float_host(dbg[6]), dbg[7], s->act_dchunks);
            }
        }
    }
#endif
    const int fin_blocks = std::max(1, std::min(2 * s->tma_ctas, (nhist * 3 + 2 + 7) / 8));   // one warp per row of the partial matrix
    if (s->comm == nullptr) {
        k_qn_fin1<<<fin_blocks, 256, 0, st>>>(nhist, s->partial, s->act_dchunks, s->coef, s->cap, s->dbuf, s->norm_part, norm_blocks, s->ctrl,
                                              s->rel_trace, s->abs_trace, n, s->eps, 1e3 * PSI_D, s->threshold);
        PSI_CK_LAUNCH();
    } else if (s->part != nullptr && s->part->p2p) {
        // mesh-partitioned over peer-mapped memory: reduction, exchange of the 3(n−1)+4 fp64 sums and the stop rules in ONE kernel
        if (3 * nhist + 4 > PSI_RED_MAX) PSI_FAIL("threshold exceeds the reduce slots of the partitioned solve");
        PartDev D = s->part->dev;
        k_qn_fin1_p2p<<<fin_blocks, 256, 0, st>>>(D, nhist, s->partial, s->act_dchunks, s->coef, s->cap, s->dbuf, s->norm_part, norm_blocks, s->ctrl,
                                                  s->rel_trace, s->abs_trace, n, s->eps, 1e3 * PSI_D, s->threshold, s->part->next(s->part->cnt_red));
        PSI_CK_LAUNCH();
    } else {
        // mesh-partitioned, NCCL fallback: local fp64 sums → ONE all-reduce of 3(n−1)+4 doubles per step → coefficients and stop rules
        k_qn_fin1_local<<<fin_blocks, 256, 0, st>>>(nhist, s->partial, s->act_dchunks, s->dbuf, s->norm_part, norm_blocks, s->ctrl);
        PSI_CK_LAUNCH();
        if (allreduce_f64(s->comm, s->dbuf, (size_t)3 * nhist + 4, st)) return -1;
        k_qn_fin1_global<<<1, 256, 0, st>>>(nhist, s->dbuf, s->coef, s->cap, s->ctrl, s->rel_trace, s->abs_trace, n, s->eps, 1e3 * PSI_D,
                                            s->threshold);
        PSI_CK_LAUNCH();
        s->launches += 1;
    }
    float* xt = s->xtrace ? s->xtrace + (int64_t)(n + 1) * s->stride : nullptr;
    if (n >= s->threshold) xt = nullptr;
    // pass 2: rank-one update, new update direction and the step itself (history read once, 4 vectors read + 4 written besides)
    prof_begin(s, n, 2, (2.0 * nhist + 8.0) * vec, st);
    k_qn_axpy_tma<<<s->tma_ctas, TMA_THREADS, axpy_tma_smem(nhist), st>>>(s->hist, nhist, n, s->coef, s->cap, s->dbuf, s->dx, s->dg, s->g, s->x,
                                                                         s->best, xt, (s->act_numel + 3) / 4, s->ctrl);
    prof_end(s, n, 2, st);
    PSI_CK_LAUNCH();
    s->launches += 3;
    return 0;
}

static int qn_poll(psi_solver* s, cudaStream_t st) {
    PSI_CK(cudaMemcpyAsync(s->h_ctrl, s->ctrl, sizeof(QnCtrl), cudaMemcpyDeviceToHost, st));
    PSI_CK(cudaStreamSynchronize(st));
    return 0;
}

static int qn_finish(psi_solver* s, float* result, psi_solve_stats_t* stats, double* rel_trace, double* abs_trace, cudaStream_t st) {
    if (qn_poll(s, st)) return -1;
    if (s->part != nullptr) {
        // the exchanges queued behind the stop were no-ops and acknowledged nothing; the all-reduce of the last executed step has
        // ordered every real one — the next exchange needs no acknowledgement
        s->part->last_halo = 0; s->part->last_sb = 0;
    }
    const QnCtrl& c = *s->h_ctrl;
    const int ran = c.nstep;
    // the last executed step stopped before its rank-one update unless it ran out of steps: its axpy/fin2 pairs are no-ops
    if (s->profile) {
        const int stopped = (c.stop_reason != 0) ? ran : ran + 1;
        for (int cls = 2; cls < PROF_CLASSES; ++cls) {
            const size_t i = ((size_t)stopped * PROF_CLASSES + cls);
            if (stopped <= s->cap + 1 && i < s->ev_bytes.size()) s->ev_bytes[i] = -1.0;
        }
        for (int step = ran + 1; step <= s->cap + 1; ++step)
            for (int cls = 0; cls < PROF_CLASSES; ++cls) {
                const size_t i = ((size_t)step * PROF_CLASSES + cls);
                if (i < s->ev_bytes.size()) s->ev_bytes[i] = -1.0;
            }
        prof_collect(s, ran);
    }
    if (result != nullptr && s->numel > 0)
        PSI_CK(cudaMemcpyAsync(result, s->best, s->numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (stats != nullptr) {
        stats->lowest = c.best_rel;
        stats->nstep = c.best_step_rel;
        stats->steps_run = ran;
        stats->prot_break = c.prot_break;
        stats->stop_reason = (c.stop_reason == 5 || c.stop_reason == 4) ? 0 : c.stop_reason;
        stats->f_evals = s->f_evals;
        stats->launches = s->launches;
    }
    if (rel_trace != nullptr || abs_trace != nullptr) {
        std::vector<double> tmp(s->threshold + 1);
        for (int which = 0; which < 2; ++which) {
            double* dst = which == 0 ? rel_trace : abs_trace;
            if (dst == nullptr) continue;
            if (ran > 0)
                PSI_CK(cudaMemcpyAsync(tmp.data(), which == 0 ? s->rel_trace : s->abs_trace, ran * sizeof(double), cudaMemcpyDeviceToHost, st));
            PSI_CK(cudaStreamSynchronize(st));
            const double low = which == 0 ? c.best_rel : c.best_abs;
            for (int i = 0; i < s->threshold + 1; ++i) dst[i] = i < ran ? tmp[i] : low;   // padded with the lowest value (solver.py:195-197)
        }
    }
    s->active = false;
    return 0;
}

static int op_eval(psi_solver* s, psi_graph* g, int kind, int op, const float* aux, float* out, cudaStream_t st) {
    SolverEpi E{s->g, s->dg, s->norm_part, &s->ctrl->done};
    const int step = s->f_evals;          // evaluation 0 yields g_0, evaluation n belongs to step n
    s->f_evals += 1;
    int rc;
    // the layer operator reads the neighbours' rows of the iterate (ghost rows from their owners); the VJP exchanges S̄ between its phases
    if (g->part != nullptr && op == PSI_OP_LAYER && halo_refresh(g, s->x, 0, &s->ctrl->done, st)) return -1;
    prof_begin(s, step, 0, s->op_bytes, st);
    if (op == PSI_OP_LAYER) {
        s->launches += fused_pre(g, kind) ? 1 : 2;
        rc = launch_layer<true>(g, kind, s->x, aux, out, E, st);
    } else {
        s->launches += 2;
        rc = launch_vjp<true>(g, kind, s->x, aux, out, E, st);
    }
    prof_end(s, step, 0, st);
    return rc;
}

// algorithmic bytes of one operator evaluation inside a solve: every array the kernel must touch once
static double operator_bytes(const psi_graph* g, int kind, int op) {
    const double N = (double)g->N, E = (double)g->E, Nd = (double)g->n_dir;
    const int lists = (kind == PSI_KIND_MIXED) ? 3 : 2;
    if (op == PSI_OP_LAYER)   // h read + tag + prb + 2 slice offsets/32 ; records 16 B per edge per list ; h0 on Dirichlet rows ; epilogue g, δg
        return N * (40.0 + 1.0 + 4.0 * g->prb_dim + 0.5) + lists * E * 16.0 + Nd * 40.0 + N * 120.0 + (kind == PSI_KIND_MIXED ? N * 8.0 : 0.0);
    // VJP: phase A 253 B read + 120 B written per node; phase B {j, mask} 8 B per edge per list + S̄ 80 + D 40 + grad 40 + y 40 + epilogue 120
    return N * (253.0 + 120.0) + 2.0 * E * 8.0 + N * (80.0 + 40.0 + 40.0 + 40.0 + 120.0);
}

extern "C" int psi_solver_broyden(psi_solver_t* s, psi_graph_t* g, int kind, int op, const float* dev_x0, const float* dev_aux,
                                  int threshold, double eps, float* dev_result, psi_solve_stats_t* stats, double* rel_trace,
                                  double* abs_trace, float* dev_xtrace, void* stream) {
    if (s == nullptr) PSI_FAIL("psi_solver_broyden: null solver");
    if (check_kind(g, kind)) return -1;
    if (op != PSI_OP_LAYER && op != PSI_OP_VJP) PSI_FAIL("psi_solver_broyden: unknown operator");
    if (op == PSI_OP_VJP && (!g->vjp_ready || g->vjp_kind != kind)) PSI_FAIL("psi_solver_broyden: call psi_vjp_prepare first");
    if (op == PSI_OP_VJP && kind != PSI_KIND_DIRICHLET && kind != PSI_KIND_MIXED) PSI_FAIL("psi_solver_broyden: no VJP for this kind");
    if (g->N * PSI_D != s->numel) PSI_FAIL("psi_solver_broyden: solver workspace size does not match the graph");
    if (g->N > 0 && dev_aux == nullptr && !(op == PSI_OP_LAYER && kind == PSI_KIND_DSS)) PSI_FAIL("psi_solver_broyden: null aux (h0 / grad)");
    cudaStream_t st = as_stream(stream);
    if (qn_begin(s, dev_x0, threshold, eps, dev_xtrace, st, (int64_t)g->dev.n_compute * PSI_D, g->part ? g->part->comm : nullptr)) return -1;
    s->part = (s->comm != nullptr) ? g->part : nullptr;
    if (s->part != nullptr) s->part->new_epoch();
    s->op_bytes = operator_bytes(g, kind, op);
    const int norm_blocks = (int)node_grid(g->dev.n_compute);
    if (g->N > 0) {
        if (op_eval(s, g, kind, op, dev_aux, nullptr, st)) return -1;        // g_0 = op(x_0) − x_0
        if (qn_first(s, st)) return -1;
        // The host runs at most AHEAD steps in front of the device (an event per step) and reads the stop word the device writes into
        // mapped host memory — no stream synchronisation, no copy; kernels of the (≤ AHEAD − 1) steps queued past the stop are no-ops.
        // A mesh-partitioned solve must leave the loop at the SAME step on every rank (the NCCL fallback enqueues real sends and
        // receives for every queued step): there the stop flag is polled at fixed steps instead.
        constexpr int AHEAD = 3;
        const bool lockstep = (s->comm != nullptr);
        for (int n = 1; n <= threshold; ++n) {
            if (!lockstep && n > AHEAD) {
                PSI_CK(cudaEventSynchronize(s->ahead[(n - AHEAD) % 4]));
                if (*reinterpret_cast<volatile int*>(s->h_done)) break;
            }
            if (op_eval(s, g, kind, op, dev_aux, nullptr, st)) return -1;
            if (qn_update(s, n, norm_blocks, st)) return -1;
            if (!lockstep) PSI_CK(cudaEventRecord(s->ahead[n % 4], st));
            else if (n % 4 == 0 && n < threshold) {
                if (qn_poll(s, st)) return -1;
                if (s->h_ctrl->done) break;
            }
        }
    }
    return qn_finish(s, dev_result, stats, rel_trace, abs_trace, st);
}

// ---- step API for an arbitrary operator -----------------------------------------------------------------
extern "C" int psi_broyden_begin(psi_solver_t* s, const float* dev_x0, int threshold, double eps, float* dev_xtrace, void* stream) {
    return qn_begin(s, dev_x0, threshold, eps, dev_xtrace, as_stream(stream));
}

extern "C" const float* psi_broyden_x(const psi_solver_t* s) { return s ? s->x : nullptr; }

static int qn_post(psi_solver* s, const float* fx, cudaStream_t st) {
    if (s->numel == 0) return 0;
    if (fx == nullptr) PSI_FAIL("null operator output");
    s->norm_blocks = (int)((s->act_numel + QN_THREADS * 4 - 1) / (QN_THREADS * 4));
    k_qn_post<<<s->norm_blocks, QN_THREADS, 0, st>>>(s->act_numel, fx, s->x, s->g, s->dg, s->norm_part, &s->ctrl->done);
    PSI_CK_LAUNCH();
    s->launches += 1;
    s->f_evals += 1;
    return 0;
}

extern "C" int psi_broyden_first(psi_solver_t* s, const float* dev_fx, void* stream) {
    if (s == nullptr || !s->active) PSI_FAIL("psi_broyden_first: call psi_broyden_begin first");
    cudaStream_t st = as_stream(stream);
    if (qn_post(s, dev_fx, st)) return -1;
    if (s->numel > 0 && qn_first(s, st)) return -1;
    return 0;
}

extern "C" int psi_broyden_step(psi_solver_t* s, const float* dev_fx, int* done, void* stream) {
    if (s == nullptr || !s->active) PSI_FAIL("psi_broyden_step: call psi_broyden_begin first");
    cudaStream_t st = as_stream(stream);
    if (s->n >= s->threshold || s->numel == 0) { if (done) *done = 1; return 0; }
    s->n += 1;
    if (qn_post(s, dev_fx, st)) return -1;
    if (qn_update(s, s->n, s->norm_blocks, st)) return -1;
    if (done != nullptr) {
        if (qn_poll(s, st)) return -1;
        *done = s->h_ctrl->done;
    }
    return 0;
}

extern "C" int psi_broyden_finish(psi_solver_t* s, float* dev_result, psi_solve_stats_t* stats, double* rel_trace, double* abs_trace,
                                  void* stream) {
    if (s == nullptr || !s->active) PSI_FAIL("psi_broyden_finish: call psi_broyden_begin first");
    return qn_finish(s, dev_result, stats, rel_trace, abs_trace, as_stream(stream));
}

// ---- teacher-forced single rank-one update (parity tests) -----------------------------------------------
__global__ void k_forced_load(int64_t numel, const float* __restrict__ x, const float* __restrict__ gx, const float* __restrict__ xn,
                              const float* __restrict__ gn, float* __restrict__ sx, float* __restrict__ sg, float* __restrict__ sdx,
                              float* __restrict__ sdg) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= numel) return;
    sx[i] = xn[i];
    sg[i] = gn[i];
    sdx[i] = xn[i] - x[i];
    sdg[i] = gn[i] - gx[i];
}

extern "C" int psi_broyden_forced_step(psi_solver_t* s, int n, const float* dev_x, const float* dev_gx, const float* dev_xnew,
                                       const float* dev_gnew, const float* dev_U, const float* dev_V, float* dev_u, float* dev_v,
                                       float* dev_update, void* stream) {
    if (s == nullptr) PSI_FAIL("psi_broyden_forced_step: null solver");
    if (n < 1 || n > s->cap) PSI_FAIL("psi_broyden_forced_step: n out of range");
    if (s->numel == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (hist_ensure(s, n - 1, st)) return -1;
    s->threshold = s->cap + 1; s->eps = 0.0; s->xtrace = nullptr; s->launches = 0; s->f_evals = 0; s->comm = nullptr;
    s->act_numel = s->numel; s->act_chunks = (int)((s->numel + QN_CHUNK - 1) / QN_CHUNK); s->act_dchunks = (int)((s->numel + DOTS_CH - 1) / DOTS_CH);
    s->last_act = -1;   // the next solve re-establishes the zero padding
    k_qn_ctrl_init<<<1, 1, 0, st>>>(s->ctrl, nullptr);
    for (int k = 0; k < n - 1; ++k) {
        float* du = s->hist.U[k / s->hist.slab_vecs] + (int64_t)(k % s->hist.slab_vecs) * s->stride;
        float* dv = s->hist.V[k / s->hist.slab_vecs] + (int64_t)(k % s->hist.slab_vecs) * s->stride;
        PSI_CK(cudaMemsetAsync(du, 0, s->stride * sizeof(float), st));
        PSI_CK(cudaMemsetAsync(dv, 0, s->stride * sizeof(float), st));
        PSI_CK(cudaMemcpyAsync(du, dev_U + (int64_t)k * s->numel, s->numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
        PSI_CK(cudaMemcpyAsync(dv, dev_V + (int64_t)k * s->numel, s->numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    k_forced_load<<<(unsigned)((s->numel + 255) / 256), 256, 0, st>>>(s->numel, dev_x, dev_gx, dev_xnew, dev_gnew, s->x, s->g, s->dx, s->dg);
    PSI_CK_LAUNCH();
    PSI_CK(cudaMemsetAsync(s->norm_part, 0, 2 * sizeof(float), st));
    const int saved_thr = s->threshold;
    if (qn_update(s, n, 1, st)) return -1;
    s->threshold = saved_thr;
    float* du = s->hist.U[(n - 1) / s->hist.slab_vecs] + (int64_t)((n - 1) % s->hist.slab_vecs) * s->stride;
    float* dv = s->hist.V[(n - 1) / s->hist.slab_vecs] + (int64_t)((n - 1) % s->hist.slab_vecs) * s->stride;
    if (dev_u) PSI_CK(cudaMemcpyAsync(dev_u, du, s->numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (dev_v) PSI_CK(cudaMemcpyAsync(dev_v, dv, s->numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (dev_update) PSI_CK(cudaMemcpyAsync(dev_update, s->dx, s->numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

// ================================================================================================
// Picard iteration (forward_iteration, solver.py:301-341) on the layer operator
// ================================================================================================
extern "C" int psi_solver_picard(psi_solver_t* s, psi_graph_t* g, int kind, const float* dev_x0, const float* dev_h0, int threshold,
                                 double eps, float* dev_result, psi_solve_stats_t* stats, double* rel_trace, double* abs_trace,
                                 float* dev_xtrace, void* stream) {
    if (s == nullptr) PSI_FAIL("psi_solver_picard: null solver");
    if (check_kind(g, kind)) return -1;
    if (g->N * PSI_D != s->numel) PSI_FAIL("psi_solver_picard: solver workspace size does not match the graph");
    if (g->part != nullptr) PSI_FAIL("psi_solver_picard: mesh-partitioned graphs are solved with psi_solver_broyden");
    if (threshold < 0 || threshold > s->cap) PSI_FAIL("psi_solver_picard: threshold exceeds the solver workspace");
    cudaStream_t st = as_stream(stream);
    if (qn_begin(s, dev_x0, threshold, eps, dev_xtrace, st)) return -1;       // xest_trace[0] = z0 (solver.py:303-304)
    const int norm_blocks = (int)node_grid(g->N);
    int it = 0, evals = 0;
    double last_rel = 0.0;
    if (g->N > 0) {
        // z = f(z_prev): the epilogue yields ‖z − z_prev‖² and ‖z‖² partials; s->fx receives z, then the buffers swap.
        const int poll = s->numel < (1 << 22) ? 16 : 4;
        bool done = false;
        while (!done) {
            SolverEpi E{s->g, s->dg, s->norm_part, &s->ctrl->done};
            if (launch_layer<true>(g, kind, s->x, dev_h0, s->fx, E, st)) return -1;
            k_picard_fin<<<1, 32, 0, st>>>(s->norm_part, norm_blocks, s->ctrl, s->rel_trace, s->abs_trace, evals, (float)eps, threshold);
            PSI_CK_LAUNCH();
            std::swap(s->x, s->fx);
            ++evals;
            s->launches += 3;
            if (dev_xtrace != nullptr && evals <= threshold + 1)                  // z_est.append(z) (rows past the stop are never read)
                PSI_CK(cudaMemcpyAsync(dev_xtrace + (size_t)evals * s->stride, s->x, s->stride * sizeof(float), cudaMemcpyDeviceToDevice, st));
            if (evals % poll == 0 || evals > threshold) {
                if (qn_poll(s, st)) return -1;
                done = s->h_ctrl->done != 0;
            }
        }
        // after the stop, later evaluations were no-ops but the host kept swapping: the live iterate is the
        // buffer written by evaluation number ctrl.nstep (1-based); recover it from the parity of the swaps.
        const int used = s->h_ctrl->nstep;                     // evaluations that actually ran
        if ((evals - used) % 2 != 0) std::swap(s->x, s->fx);
        it = used - 1;
        last_rel = s->h_ctrl->best_rel;
        evals = used;
    }
    if (dev_result != nullptr && s->numel > 0)
        PSI_CK(cudaMemcpyAsync(dev_result, s->x, s->numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (stats != nullptr) {
        stats->lowest = last_rel; stats->nstep = it; stats->steps_run = it; stats->prot_break = 0;
        stats->stop_reason = (it < threshold) ? 1 : 0; stats->f_evals = evals; stats->launches = s->launches;
    }
    for (int which = 0; which < 2; ++which) {
        double* dst = which == 0 ? rel_trace : abs_trace;
        if (dst == nullptr) continue;
        for (int i = 0; i < threshold + 1; ++i) dst[i] = std::numeric_limits<double>::quiet_NaN();
        if (evals > 0) PSI_CK(cudaMemcpyAsync(dst, which == 0 ? s->rel_trace : s->abs_trace, evals * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    PSI_CK(cudaStreamSynchronize(st));
    s->active = false;
    return 0;
}

// ================================================================================================
// Anderson acceleration (solver.py:215-293): one state machine serves the fused loop on the layer operator and the step API
// for an arbitrary operator.  Evaluation e (0-based) reads X[e % m] and writes F[e % m]:
//   e = 0: X[0] = x0 ; e = 1: X[1] = F[0] ; e = k ≥ 2: X[k % m] = β·Σ α_i F_i + (1−β)·Σ α_i X_i with α from the bordered system.
// ================================================================================================
static int and_ensure(psi_solver* s, int m) {
    if (s->and_m >= m) return 0;
    if (s->and_X) { cudaFree(s->and_X); cudaFree(s->and_F); cudaFree(s->and_small); s->and_X = s->and_F = s->and_small = nullptr; }
    if (solver_alloc(s, (void**)&s->and_X, (size_t)m * s->stride * sizeof(float))) return -1;
    if (solver_alloc(s, (void**)&s->and_F, (size_t)m * s->stride * sizeof(float))) return -1;
    if (solver_alloc(s, (void**)&s->and_small, (size_t)(AND_MAX_M * AND_MAX_M * (s->num_chunks + 1) + 64) * sizeof(float))) return -1;
    s->and_m = m;
    return 0;
}

static int and_begin(psi_solver* s, const float* x0, int m, double lam, int threshold, double eps, double beta, float* xtrace, cudaStream_t st) {
    if (m < 2 || m > AND_MAX_M) PSI_FAIL("anderson: m must be in [2, 8]");
    if (threshold < 0 || threshold > s->cap) PSI_FAIL("anderson: threshold exceeds the solver workspace");
    if (and_ensure(s, m)) return -1;
    if (qn_begin(s, x0, threshold, eps, nullptr, st)) return -1;
    s->xtrace = xtrace;
    s->and_cur_m = m; s->and_lam = lam; s->and_beta = beta; s->and_e = 0;
    if (s->numel > 0) {
        const size_t vb = s->stride * sizeof(float);
        PSI_CK(cudaMemsetAsync(s->and_X, 0, (size_t)m * vb, st));
        PSI_CK(cudaMemsetAsync(s->and_F, 0, (size_t)m * vb, st));
        PSI_CK(cudaMemcpyAsync(s->and_X, x0, s->numel * sizeof(float), cudaMemcpyDeviceToDevice, st));        // X[0] = x0 (solver.py:227)
        if (xtrace != nullptr) PSI_CK(cudaMemcpyAsync(xtrace, s->and_X, vb, cudaMemcpyDeviceToDevice, st));     // xest_trace[0] = x0 (:243)
    }
    return 0;
}
static inline float* and_in(psi_solver* s) { return s->and_X + (size_t)(s->and_e % s->and_cur_m) * s->stride; }
static inline float* and_out(psi_solver* s) { return s->and_F + (size_t)(s->and_e % s->and_cur_m) * s->stride; }
static inline int and_total_evals(const psi_solver* s) { return std::max(s->threshold, 2); }

// bookkeeping after evaluation number s->and_e has been written to and_out(s)
static int and_advance(psi_solver* s, cudaStream_t st) {
    const int e = s->and_e, m = s->and_cur_m;
    const size_t vb = s->stride * sizeof(float);
    float* xs = and_in(s);
    float* fs = and_out(s);
    float* part = s->and_small;                                                           // [n*n][num_chunks]
    float* alpha = s->and_small + (size_t)AND_MAX_M * AND_MAX_M * s->num_chunks;          // [AND_MAX_M]
    s->f_evals += 1;
    if (e == 0) {
        PSI_CK(cudaMemcpyAsync(s->and_X + s->stride, s->and_F, vb, cudaMemcpyDeviceToDevice, st));             // X[1] = F[0] (:229)
    } else if (e >= 2) {
        k_and_post<<<s->num_chunks, QN_THREADS, 0, st>>>(xs, fs, s->best, s->norm_part, s->num_chunks, &s->ctrl->done);
        k_and_fin<<<1, 32, 0, st>>>(s->norm_part, s->num_chunks, s->ctrl, s->rel_trace, s->abs_trace, e, s->eps);
        k_and_keep<<<s->axpy_ctas, QN_THREADS, 0, st>>>(xs, s->best, s->num_chunks, s->ctrl, e);
        PSI_CK_LAUNCH();
        s->launches += 3;
        if (s->xtrace != nullptr)                                                          // xest_trace.append(lowest_xest) (:273)
            PSI_CK(cudaMemcpyAsync(s->xtrace + (size_t)(e - 1) * s->stride, s->best, vb, cudaMemcpyDeviceToDevice, st));
    }
    s->and_e = e + 1;
    const int k = s->and_e;
    if (k >= 2 && k < s->threshold) {
        const int n = std::min(k, m);
        k_and_gram<<<s->num_chunks, QN_THREADS, 0, st>>>(s->and_X, s->and_F, s->stride, n, part, s->num_chunks, &s->ctrl->done);
        k_and_solve<<<1, 32, 0, st>>>(part, s->num_chunks, n, (float)s->and_lam, alpha, &s->ctrl->done);
        k_and_mix<<<s->axpy_ctas, QN_THREADS, 0, st>>>(s->and_X, s->and_F, s->stride, n, k % m, alpha, (float)s->and_beta, s->num_chunks, &s->ctrl->done);
        PSI_CK_LAUNCH();
        s->launches += 3;
    }
    return 0;
}

static int and_finish(psi_solver* s, float* dev_result, psi_solve_stats_t* stats, double* rel_trace, double* abs_trace, cudaStream_t st) {
    if (qn_poll(s, st)) return -1;
    const QnCtrl& c = *s->h_ctrl;
    const int threshold = s->threshold;
    const int ran = c.nstep >= 2 ? c.nstep - 1 : 0;       // trace entries written (k = 2 .. nstep)
    if (dev_result != nullptr && s->numel > 0)
        PSI_CK(cudaMemcpyAsync(dev_result, s->best, s->numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (stats != nullptr) {
        stats->lowest = c.best_rel; stats->nstep = c.best_step_rel; stats->steps_run = ran; stats->prot_break = 0;
        stats->stop_reason = c.done ? 1 : 0; stats->f_evals = s->f_evals; stats->launches = s->launches;
    }
    // trace length: threshold−2 entries (k = 2..threshold−1), padded with the lowest after an early stop (solver.py:275-278)
    const int len = std::max(threshold - 2, 0);
    std::vector<double> tmp(std::max(ran, 1));
    for (int which = 0; which < 2; ++which) {
        double* dst = which == 0 ? rel_trace : abs_trace;
        if (dst == nullptr) continue;
        if (ran > 0) PSI_CK(cudaMemcpyAsync(tmp.data(), which == 0 ? s->rel_trace : s->abs_trace, ran * sizeof(double), cudaMemcpyDeviceToHost, st));
        PSI_CK(cudaStreamSynchronize(st));
        const double low = which == 0 ? c.best_rel : c.best_abs;
        for (int i = 0; i < threshold + 1; ++i) dst[i] = i < ran ? tmp[i] : (i < len ? low : std::numeric_limits<double>::quiet_NaN());
    }
    PSI_CK(cudaStreamSynchronize(st));
    s->active = false;
    return 0;
}

extern "C" int psi_solver_anderson(psi_solver_t* s, psi_graph_t* g, int kind, const float* dev_x0, const float* dev_h0, int m, double lam,
                                   int threshold, double eps, double beta, float* dev_result, psi_solve_stats_t* stats,
                                   double* rel_trace, double* abs_trace, float* dev_xtrace, void* stream) {
    if (s == nullptr) PSI_FAIL("psi_solver_anderson: null solver");
    if (check_kind(g, kind)) return -1;
    if (g->N * PSI_D != s->numel) PSI_FAIL("psi_solver_anderson: solver workspace size does not match the graph");
    if (g->part != nullptr) PSI_FAIL("psi_solver_anderson: mesh-partitioned graphs are solved with psi_solver_broyden");
    cudaStream_t st = as_stream(stream);
    if (and_begin(s, dev_x0, m, lam, threshold, eps, beta, dev_xtrace, st)) return -1;
    if (g->N > 0) {
        const SolverEpi noE{nullptr, nullptr, nullptr, nullptr};
        const int poll = s->numel < (1 << 22) ? 8 : 2;
        while (s->and_e < and_total_evals(s)) {
            if (launch_layer<false>(g, kind, and_in(s), dev_h0, and_out(s), noE, st)) return -1;   // a no-op write target once done: guarded by the kernels below
            s->launches += 2;
            if (and_advance(s, st)) return -1;
            if (s->and_e > 2 && (s->and_e - 2) % poll == 0) {
                if (qn_poll(s, st)) return -1;
                if (s->h_ctrl->done) break;
            }
        }
    }
    return and_finish(s, dev_result, stats, rel_trace, abs_trace, st);
}

// ---- teacher-forced single Anderson update (parity tests): window (X, F) of n vectors in, X[slot] = β·Σα_iF_i + (1−β)·Σα_iX_i and α out
extern "C" int psi_anderson_forced_step(psi_solver_t* s, int m, int n, int slot, double lam, double beta, const float* dev_X, const float* dev_F,
                                        float* dev_xnew, float* dev_alpha, void* stream) {
    if (s == nullptr) PSI_FAIL("psi_anderson_forced_step: null solver");
    if (m < 2 || m > AND_MAX_M || n < 1 || n > m || slot < 0 || slot >= m) PSI_FAIL("psi_anderson_forced_step: bad window");
    if (s->numel == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (and_ensure(s, m)) return -1;
    const size_t vb = s->stride * sizeof(float);
    PSI_CK(cudaMemsetAsync(s->and_X, 0, (size_t)m * vb, st));
    PSI_CK(cudaMemsetAsync(s->and_F, 0, (size_t)m * vb, st));
    for (int i = 0; i < n; ++i) {
        PSI_CK(cudaMemcpyAsync(s->and_X + (size_t)i * s->stride, dev_X + (size_t)i * s->numel, s->numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
        PSI_CK(cudaMemcpyAsync(s->and_F + (size_t)i * s->stride, dev_F + (size_t)i * s->numel, s->numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    k_qn_ctrl_init<<<1, 1, 0, st>>>(s->ctrl, nullptr);
    float* part = s->and_small;
    float* alpha = s->and_small + (size_t)AND_MAX_M * AND_MAX_M * s->num_chunks;
    k_and_gram<<<s->num_chunks, QN_THREADS, 0, st>>>(s->and_X, s->and_F, s->stride, n, part, s->num_chunks, &s->ctrl->done);
    k_and_solve<<<1, 32, 0, st>>>(part, s->num_chunks, n, (float)lam, alpha, &s->ctrl->done);
    k_and_mix<<<s->axpy_ctas, QN_THREADS, 0, st>>>(s->and_X, s->and_F, s->stride, n, slot, alpha, (float)beta, s->num_chunks, &s->ctrl->done);
    PSI_CK_LAUNCH();
    if (dev_xnew) PSI_CK(cudaMemcpyAsync(dev_xnew, s->and_X + (size_t)slot * s->stride, s->numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (dev_alpha) PSI_CK(cudaMemcpyAsync(dev_alpha, alpha, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    s->last_act = -1;
    return 0;
}

// ---- step API (arbitrary operator): begin ; loop { fx = f(psi_anderson_x()) ; psi_anderson_feed(fx) -> done? } ; finish ----------
extern "C" int psi_anderson_begin(psi_solver_t* s, const float* dev_x0, int m, double lam, int threshold, double eps, double beta,
                                  float* dev_xtrace, void* stream) {
    if (s == nullptr) PSI_FAIL("psi_anderson_begin: null solver");
    return and_begin(s, dev_x0, m, lam, threshold, eps, beta, dev_xtrace, as_stream(stream));
}
extern "C" const float* psi_anderson_x(const psi_solver_t* s) {
    return (s && s->and_X) ? s->and_X + (size_t)(s->and_e % s->and_cur_m) * s->stride : nullptr;
}
extern "C" int psi_anderson_feed(psi_solver_t* s, const float* dev_fx, int* done, void* stream) {
    if (s == nullptr || !s->active) PSI_FAIL("psi_anderson_feed: call psi_anderson_begin first");
    cudaStream_t st = as_stream(stream);
    if (s->numel == 0 || s->and_e >= and_total_evals(s)) { if (done) *done = 1; return 0; }
    if (dev_fx == nullptr) PSI_FAIL("psi_anderson_feed: null operator output");
    PSI_CK(cudaMemcpyAsync(and_out(s), dev_fx, s->numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (and_advance(s, st)) return -1;
    if (done != nullptr) {
        *done = 0;
        if (s->and_e > 2) { if (qn_poll(s, st)) return -1; *done = s->h_ctrl->done; }
        if (s->and_e >= and_total_evals(s)) *done = 1;
    }
    return 0;
}
extern "C" int psi_anderson_finish(psi_solver_t* s, float* dev_result, psi_solve_stats_t* stats, double* rel_trace, double* abs_trace, void* stream) {
    if (s == nullptr || !s->active) PSI_FAIL("psi_anderson_finish: call psi_anderson_begin first");
    return and_finish(s, dev_result, stats, rel_trace, abs_trace, as_stream(stream));
}

// ---- Picard step API (arbitrary operator): begin ; loop { fx = f(psi_picard_x()) ; psi_picard_feed(fx) -> done? } ; finish ------
extern "C" int psi_picard_begin(psi_solver_t* s, const float* dev_z0, int threshold, double eps, float* dev_xtrace, void* stream) {
    if (s == nullptr) PSI_FAIL("psi_picard_begin: null solver");
    if (threshold < 0 || threshold > s->cap) PSI_FAIL("psi_picard_begin: threshold exceeds the solver workspace");
    cudaStream_t st = as_stream(stream);
    if (qn_begin(s, dev_z0, threshold, eps, dev_xtrace, st)) return -1;       // xtrace row 0 = z0
    s->and_e = 0;
    return 0;
}
extern "C" const float* psi_picard_x(const psi_solver_t* s) { return s ? s->x : nullptr; }
extern "C" int psi_picard_feed(psi_solver_t* s, const float* dev_fx, int* done, void* stream) {
    if (s == nullptr || !s->active) PSI_FAIL("psi_picard_feed: call psi_picard_begin first");
    cudaStream_t st = as_stream(stream);
    if (s->numel == 0) { if (done) *done = 1; return 0; }
    if (dev_fx == nullptr) PSI_FAIL("psi_picard_feed: null operator output");
    // ‖z − z_prev‖², ‖z‖² partials (z = fx, z_prev = x), stop rule, then z becomes the iterate
    const int nb = (int)((s->numel + QN_THREADS * 4 - 1) / (QN_THREADS * 4));
    k_qn_post<<<nb, QN_THREADS, 0, st>>>(s->numel, dev_fx, s->x, s->g, s->dg, s->norm_part, &s->ctrl->done);
    k_picard_fin<<<1, 32, 0, st>>>(s->norm_part, nb, s->ctrl, s->rel_trace, s->abs_trace, s->and_e, (float)s->eps, s->threshold);
    PSI_CK_LAUNCH();
    PSI_CK(cudaMemcpyAsync(s->x, dev_fx, s->numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (s->xtrace != nullptr)
        PSI_CK(cudaMemcpyAsync(s->xtrace + (size_t)(s->and_e + 1) * s->stride, dev_fx, s->numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
    s->and_e += 1; s->f_evals += 1; s->launches += 2;
    if (qn_poll(s, st)) return -1;
    if (done != nullptr) *done = s->h_ctrl->done;
    return 0;
}
extern "C" int psi_picard_finish(psi_solver_t* s, float* dev_result, psi_solve_stats_t* stats, double* rel_trace, double* abs_trace, void* stream) {
    if (s == nullptr || !s->active) PSI_FAIL("psi_picard_finish: call psi_picard_begin first");
    cudaStream_t st = as_stream(stream);
    if (qn_poll(s, st)) return -1;
    const int evals = s->h_ctrl->nstep, threshold = s->threshold;
    if (dev_result != nullptr && s->numel > 0) PSI_CK(cudaMemcpyAsync(dev_result, s->x, s->numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (stats != nullptr) {
        stats->lowest = s->h_ctrl->best_rel; stats->nstep = std::max(evals - 1, 0); stats->steps_run = std::max(evals - 1, 0); stats->prot_break = 0;
        stats->stop_reason = (evals - 1 < threshold) ? 1 : 0; stats->f_evals = evals; stats->launches = s->launches;
    }
    for (int which = 0; which < 2; ++which) {
        double* dst = which == 0 ? rel_trace : abs_trace;
        if (dst == nullptr) continue;
        for (int i = 0; i < threshold + 1; ++i) dst[i] = std::numeric_limits<double>::quiet_NaN();
        if (evals > 0) PSI_CK(cudaMemcpyAsync(dst, which == 0 ? s->rel_trace : s->abs_trace, std::min(evals, threshold + 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    PSI_CK(cudaStreamSynchronize(st));
    s->active = false;
    return 0;
}
