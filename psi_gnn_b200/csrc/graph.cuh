// graph.cuh — one-off re-layout of a PyG-style batch into destination-sorted, warp-sliced CSR.
//
// Replaces what the reference redoes on every f-evaluation: remove_self_loops (model.py:342,360),
// where(tags==1) (model.py:281) and the SparseTensor construction/sort (model.py:159-163).
//
// Layout ("SELL-32", a destination-sorted CSR whose rows are stored column-major inside slices of
// 32 consecutive nodes): node i belongs to slice i>>5, lane i&31; the t-th incoming edge of node i
// lives at  recs[slice_off[i>>5] + t*32 + (i&31)].  A warp that owns one slice therefore reads
// its edge records with perfectly coalesced 512-byte requests, and each destination's sum is
// accumulated by one thread in CSR order: deterministic, no atomics.  Slices are padded to the
// longest row in the slice with j = -1 records (meshes have near-uniform degree, padding ~10 %).
//
// Four lists are built:
//   T  : off-diagonal edges grouped by col  (Phi_to   aggregates at edge_index[1]), neighbour = row
//   F  : off-diagonal edges grouped by row  (Phi_from aggregates at edge_index[0]), neighbour = col
//   Ar : all nnz grouped by row, record {col, a_ij}   (residual  A u)
//   Ac : all nnz grouped by col, record {row, a_ij}   (transpose Aᵀ v for the residual backward)
#pragma once
#include "common.cuh"
#include <cub/cub.cuh>

struct SellDev {
    const int4*    recs;        // {j, attr0, attr1, attr2} (attrs as float bits) — message lists
    const int2*    recs2;       // {j, a_ij bits}                                 — matrix lists
    const int64_t* slice_off;   // [num_slices+1], in records
    uint32_t*      xmask;       // per-slot {neighbour, cross ReLU mask} pairs (int2) for the VJP (message lists only)
};

struct GraphDev {
    int      N;
    int      n_compute;         // rows the operator kernels produce: N, or the owned rows of a mesh partition (ghost rows follow)
    int      num_slices;
    int      prb_dim;
    SellDev  T, F, Ar, Ac;
    const uint8_t* tag;         // bit0 = Dirichlet, bit1 = Neumann
    const float*   prb;         // [N, prb_dim]
    const float*   nrm;         // [N, 2] or nullptr
};

// cached linearisation point for the VJP (filled by psi_vjp_prepare)
struct VjpCacheDev {
    float* rhat;    // [N,10] normalised pre-LN residual
    float* rstd;    // [N]
    float* alpha;   // [N]   gate
    float* m;       // [N,10] update MLP output
    float* cnt;     // [N,3,10] ReLU-activity counts of the to / from / neumann edge MLPs at the destination
    uint32_t* nmask; // [N] bits 0-9 update hidden ReLU mask, bits 10-19 neumann-update hidden mask
    float* Sb;      // [N,2,10] phase-A output: W2ᵀ·m̄p for the 'to' list and for the 'from'/'neumann' list
    float* Dloc;    // [N,10]   phase-A output: node-local part of Jᵀy
};

struct psi_graph {
    int64_t N = 0, nnz = 0, E = 0, n_dir = 0, n_neu = 0;
    int attr_dim = 3, prb_dim = 2, tag_dim = 1;
    int64_t slots_T = 0, slots_F = 0, slots_Ar = 0, slots_Ac = 0;
    int64_t bytes = 0;
    GraphDev dev{};
    VjpCacheDev vjp{};
    bool vjp_ready = false;
    int vjp_kind = -1;
    // owned allocations
    void* p_recs_T = nullptr; void* p_recs_F = nullptr; void* p_recs_Ar = nullptr; void* p_recs_Ac = nullptr;
    void* p_off_T = nullptr; void* p_off_F = nullptr; void* p_off_Ar = nullptr; void* p_off_Ac = nullptr;
    void* p_xm_T = nullptr; void* p_xm_F = nullptr;
    void* p_tag = nullptr; void* p_prb = nullptr; void* p_nrm = nullptr;
    void* p_vjp = nullptr;
    float* p_q = nullptr;         // [3][N][10] pre-pass scratch of the layer kernel (W1j·h per edge MLP), allocated on first use
    float* p_hstar = nullptr;     // [N][10] private copy of the linearisation point with refreshed ghost rows (mesh-partitioned VJP)
    float* p_scratch = nullptr;   // small per-graph scratch for residual partial sums
    struct Partition* part = nullptr;   // set by psi_graph_set_partition (mesh-partitioned solve)
    int64_t scratch_floats = 0;
};

// ------------------------------------------------------------------------------------------------
// build kernels
// ------------------------------------------------------------------------------------------------

// key = grouping node (or N for dropped self loops), value = edge id
// An entry whose row or column lies outside [0, N) is dropped like a self loop (key N) and counted in *bad: the host
// fails the build with "edge index out of range" (PyG / torch_sparse raise an index error there; gathering h[nb] would
// otherwise read out of bounds).
__global__ void k_graph_keys(int64_t nnz, int N, const int64_t* __restrict__ ei, int by_col, int drop_diag,
                             int* __restrict__ keys, int* __restrict__ vals, int* __restrict__ bad) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const int64_t r64 = ei[e], c64 = ei[nnz + e];
    if (r64 < 0 || r64 >= N || c64 < 0 || c64 >= N) {
        atomicAdd(bad, 1);
        keys[e] = N;
        vals[e] = (int)e;
        return;
    }
    int r = (int)r64, c = (int)c64;
    int k = by_col ? c : r;
    if (drop_diag && r == c) k = N;
    keys[e] = k;
    vals[e] = (int)e;
}

// ptr[i] = first position in sorted keys with key >= i, for i in [0, N]
__global__ void k_graph_ptr(int64_t nnz, int N, const int* __restrict__ skeys, int* __restrict__ ptr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > N) return;
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (skeys[mid] < i) lo = mid + 1; else hi = mid;
    }
    ptr[i] = (int)lo;
}

// one warp per slice: width = max degree over its 32 nodes; out = 32*width (records in the slice)
__global__ void k_graph_slice_width(int N, int num_slices, const int* __restrict__ ptr, int64_t* __restrict__ slice_recs) {
    int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (gw >= num_slices) return;
    int i = gw * 32 + lane;
    int deg = (i < N) ? (ptr[i + 1] - ptr[i]) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) deg = max(deg, __shfl_xor_sync(0xffffffffu, deg, o));
    if (lane == 0) slice_recs[gw] = (int64_t)deg * 32;
}

__global__ void k_graph_fill_msg(int N, int64_t nnz, const int64_t* __restrict__ ei, const float* __restrict__ attr,
                                 int attr_dim, int by_col, const int* __restrict__ ptr, const int* __restrict__ svals,
                                 const int64_t* __restrict__ slice_off, int4* __restrict__ recs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int64_t base = slice_off[i >> 5] + (i & 31);
    int p0 = ptr[i], p1 = ptr[i + 1];
    for (int p = p0; p < p1; ++p) {
        int e = svals[p];
        int nb = by_col ? (int)ei[e] : (int)ei[nnz + e];
        float a0 = attr[(int64_t)e * attr_dim];
        float a1 = attr_dim > 1 ? attr[(int64_t)e * attr_dim + 1] : 0.f;
        float a2 = attr_dim > 2 ? attr[(int64_t)e * attr_dim + 2] : 0.f;
        recs[base + (int64_t)(p - p0) * 32] = make_int4(nb, __float_as_int(a0), __float_as_int(a1), __float_as_int(a2));
    }
}

__global__ void k_graph_fill_mat(int N, int64_t nnz, const int64_t* __restrict__ ei, const float* __restrict__ aij,
                                 int by_col, const int* __restrict__ ptr, const int* __restrict__ svals,
                                 const int64_t* __restrict__ slice_off, int2* __restrict__ recs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int64_t base = slice_off[i >> 5] + (i & 31);
    int p0 = ptr[i], p1 = ptr[i + 1];
    for (int p = p0; p < p1; ++p) {
        int e = svals[p];
        int nb = by_col ? (int)ei[e] : (int)ei[nnz + e];
        recs[base + (int64_t)(p - p0) * 32] = make_int2(nb, __float_as_int(aij[e]));
    }
}

// tags -> byte mask; counts[0] += #dirichlet, counts[1] += #neumann
__global__ void k_graph_tags(int N, const float* __restrict__ tags, int tag_dim, uint8_t* __restrict__ out,
                             unsigned long long* __restrict__ counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    uint8_t t = 0;
    if (i < N && tags != nullptr) {
        if (tag_dim == 1) {
            t = (tags[i] == 1.0f) ? 1 : 0;                       // where(batch.tags == 1), dirichlet model.py:281
        } else {
            if (tags[(int64_t)i * tag_dim + 1] == 1.0f) t |= 1;   // tags[:,1] == 1, mixed model.py:218
            if (tags[(int64_t)i * tag_dim + 2] == 1.0f) t |= 2;   // tags[:,2] == 1, mixed model.py:219
        }
    }
    if (i < N) out[i] = t;
    unsigned d = __ballot_sync(0xffffffffu, t & 1), n = __ballot_sync(0xffffffffu, t & 2);
    if ((threadIdx.x & 31) == 0) {
        if (d) atomicAdd(&counts[0], (unsigned long long)__popc(d));
        if (n) atomicAdd(&counts[1], (unsigned long long)__popc(n));
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct SellBuild {
    void* recs = nullptr; void* off = nullptr; int64_t slots = 0; int64_t kept = 0;
};

// stream-ordered temporaries that are released on every exit path of build_sell
struct TmpAllocs {
    cudaStream_t st;
    void* p[16];
    int n = 0;
    explicit TmpAllocs(cudaStream_t s) : st(s) {}
    cudaError_t get(void** out, size_t bytes) {
        cudaError_t e = psi_malloc_async(out, bytes, st);
        if (e == cudaSuccess && n < 16) p[n++] = *out;
        return e;
    }
    void keep(void* q) {          // hand ownership to the caller
        for (int i = 0; i < n; ++i)
            if (p[i] == q) p[i] = nullptr;
    }
    ~TmpAllocs() {
        for (int i = 0; i < n; ++i)
            if (p[i]) psi_free_async(p[i], st);
    }
};

// Builds one list.  msg=true: int4 records with attrs (self loops dropped); msg=false: int2 {j, a_ij}.
static int build_sell(int64_t N, int64_t nnz, const int64_t* ei, const float* attr, int attr_dim, const float* aij,
                      bool msg, bool by_col, cudaStream_t st, SellBuild* out) {
    const int num_slices = (int)((N + 31) / 32);
    int *keys = nullptr, *vals = nullptr, *skeys = nullptr, *svals = nullptr, *ptr = nullptr;
    int64_t *slice_recs = nullptr, *slice_off = nullptr;
    int* bad = nullptr;
    void* tmp = nullptr;
    void* recs = nullptr;
    size_t tmp_bytes = 0, tmp2 = 0;
    const int64_t nn = nnz > 0 ? nnz : 1;
    TmpAllocs T(st);
    PSI_CK(T.get((void**)&keys, nn * sizeof(int)));
    PSI_CK(T.get((void**)&vals, nn * sizeof(int)));
    PSI_CK(T.get((void**)&skeys, nn * sizeof(int)));
    PSI_CK(T.get((void**)&svals, nn * sizeof(int)));
    PSI_CK(T.get((void**)&ptr, (N + 2) * sizeof(int)));
    PSI_CK(T.get((void**)&slice_recs, (num_slices + 1) * sizeof(int64_t)));
    PSI_CK(T.get((void**)&slice_off, (num_slices + 1) * sizeof(int64_t)));
    PSI_CK(T.get((void**)&bad, sizeof(int)));
    PSI_CK(cudaMemsetAsync(bad, 0, sizeof(int), st));
    int end_bit = 1;
    while ((1ll << end_bit) <= N) ++end_bit;   // keys in [0, N]
    PSI_CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, skeys, vals, svals, (int)nnz, 0, end_bit, st));
    PSI_CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, slice_recs, slice_off, num_slices + 1, st));
    if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
    PSI_CK(T.get(&tmp, tmp_bytes));
    if (nnz > 0) {
        k_graph_keys<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(nnz, (int)N, ei, by_col ? 1 : 0, msg ? 1 : 0, keys, vals, bad);
        PSI_CK_LAUNCH();
        PSI_CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, skeys, vals, svals, (int)nnz, 0, end_bit, st));
    }
    k_graph_ptr<<<(unsigned)((N + 1 + 255) / 256), 256, 0, st>>>(nnz, (int)N, skeys, ptr);
    PSI_CK_LAUNCH();
    PSI_CK(cudaMemsetAsync(slice_recs, 0, (num_slices + 1) * sizeof(int64_t), st));
    if (num_slices > 0) {
        k_graph_slice_width<<<(unsigned)((num_slices * 32 + 255) / 256), 256, 0, st>>>((int)N, num_slices, ptr, slice_recs);
        PSI_CK_LAUNCH();
    }
    PSI_CK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, slice_recs, slice_off, num_slices + 1, st));
    int64_t total = 0;
    int kept = 0, n_bad = 0;
    PSI_CK(cudaMemcpyAsync(&total, slice_off + num_slices, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PSI_CK(cudaMemcpyAsync(&kept, ptr + N, sizeof(int), cudaMemcpyDeviceToHost, st));
    PSI_CK(cudaMemcpyAsync(&n_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    PSI_CK(cudaStreamSynchronize(st));
    if (n_bad > 0) PSI_FAIL("psi_graph_create: edge index out of range (" + std::to_string(n_bad) + " entries outside [0, num_nodes))");
    const size_t rec_bytes = msg ? sizeof(int4) : sizeof(int2);
    PSI_CK(T.get(&recs, (total > 0 ? total : 1) * rec_bytes));
    PSI_CK(cudaMemsetAsync(recs, 0xFF, (total > 0 ? total : 1) * rec_bytes, st));   // j = -1 everywhere
    if (N > 0 && nnz > 0) {
        if (msg)
            k_graph_fill_msg<<<(unsigned)((N + 127) / 128), 128, 0, st>>>((int)N, nnz, ei, attr, attr_dim, by_col ? 1 : 0, ptr,
                                                                         svals, slice_off, (int4*)recs);
        else
            k_graph_fill_mat<<<(unsigned)((N + 127) / 128), 128, 0, st>>>((int)N, nnz, ei, aij, by_col ? 1 : 0, ptr, svals,
                                                                         slice_off, (int2*)recs);
        PSI_CK_LAUNCH();
    }
    T.keep(recs);
    T.keep(slice_off);
    out->recs = recs; out->off = slice_off; out->slots = total; out->kept = kept;
    return 0;
}
