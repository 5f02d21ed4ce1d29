// layer.cuh — the fused message-passing layer f_theta: a per-node pre-pass + one fused kernel per application.
//
//   k_layer_pre      Q_w[j] = W1j_w · h_j for the 2 (3) edge MLPs, once per node — the first edge layer is linear in the
//                    concatenation [h_i, h_j, a], so its h_j block is a per-SOURCE quantity: the per-edge work
//                    drops from 130 to 30 FMAs + 10 adds.
//   k_layer_forward  one thread owns one destination node (one warp = one 32-node slice of the SELL lists):
//     1. coalesced 16-byte edge records {j, a0, a1, a2}; the 32 neighbour rows Q[j] of a trip are gathered COOPERATIVELY: lane l
//        loads float2 pieces l, l+32, … of the 32·40 B the warp needs (consecutive pieces of consecutive rows → a few cache lines
//        per load instead of 32), stages them in shared memory and reads its own row back with two LDS.128 + one LDS.64.  The
//        gather, not the FMAs, bounds this kernel (ncu: L1 wavefronts, profiles/r02_a_operator.md); rows of trip t+1 are in
//        flight while trip t is consumed.
//     2. z_e = (P_i + Q_j) + W1a·a_e with P_i = b1 + W1i·h_i hoisted per destination; ReLU; summed per destination in CSR order
//        (deterministic, no atomics); the second edge layer is applied once to the sum:
//        Σ_e (W2·relu(z_e) + b2) = W2·Σ_e relu(z_e) + deg·b2
//     3. node update Psi (gate · MLP), LayerNorm, boundary masks (Dirichlet clamp, Neumann overwrite)
//     4. optional solver epilogue: g = f(x) − x, δg, and the two stopping norms
// All weights are constant-bank FFMA operands (weights.cuh).
//
// Reference semantics: dirichlet/psignn/model.py:279-300 (+ :334-368), mixed/psignn/model.py:216-245,
// dirichlet/dss/model.py:113-121, dirichlet/dsgps/model.py:143-163, mixed/dsgps/model.py:76-97.
#pragma once
#include "common.cuh"
#include "weights.cuh"
#include "graph.cuh"

enum { KIND_DIRICHLET = 0, KIND_MIXED = 1, KIND_DSS = 2, KIND_DSGPS = 3, KIND_DSGPS_MIXED = 4 };

template <int KIND> struct KindTraits {
    static constexpr bool has_neumann = (KIND == KIND_MIXED || KIND == KIND_DSGPS_MIXED);
    static constexpr int NQ = has_neumann ? 3 : 2;                 // edge MLPs whose W1j·h is pre-computed per node
    static constexpr int ATTR = (KIND == KIND_DSS) ? 1 : 3;
    static constexpr int PRB = (KIND == KIND_DIRICHLET || KIND == KIND_DSGPS) ? 2 : 3;
    static constexpr bool clamp = (KIND != KIND_DSS);              // Dirichlet rows copied from h0
};

// ---- canonical rounding of the first edge layer -------------------------------------------------------------------
// Every kernel that needs z (forward, VJP-prepare own and cross masks, parameter gradients) uses exactly these three chains, so
// that a ReLU decision is the same wherever it is taken:
//   P[o] = b1[o] + Σ_i W1i[o][i]·h_dst[i]          (fma chain from the bias)
//   Q[o] =          Σ_i W1j[o][i]·h_src[i]          (fma chain from zero)
//   z[o] = fma(W1a[o][2], a2, fma(W1a[o][1], a1, fma(W1a[o][0], a0, P[o] + Q[o])))
template <int WHICH>
__device__ __forceinline__ void edge_pre(const float (&hd)[PSI_D], float (&P)[PSI_D]) {
    const EdgeMLP& W = edge_mlp<WHICH>();
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float z = W.b1[o];
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) z = fmaf(W.W1i[o][i], hd[i], z);
        P[o] = z;
    }
}
template <int WHICH>
__device__ __forceinline__ void edge_q(const float (&hs)[PSI_D], float (&Q)[PSI_D]) {
    const EdgeMLP& W = edge_mlp<WHICH>();
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float z = 0.f;
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) z = fmaf(W.W1j[o][i], hs[i], z);
        Q[o] = z;
    }
}
template <int WHICH, int ATTR>
__device__ __forceinline__ void edge_z(const float (&P)[PSI_D], const float (&Q)[PSI_D], const int4& rec, float (&z)[PSI_D]) {
    const EdgeMLP& W = edge_mlp<WHICH>();
    const float a0 = __int_as_float(rec.y), a1 = __int_as_float(rec.z), a2 = __int_as_float(rec.w);
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float t = P[o] + Q[o];
        t = fmaf(W.W1a[o][0], a0, t);
        if (ATTR > 1) t = fmaf(W.W1a[o][1], a1, t);
        if (ATTR > 2) t = fmaf(W.W1a[o][2], a2, t);
        z[o] = t;
    }
}
// second edge layer on the aggregated hidden sums: mp = W2·S + deg·b2
template <int WHICH>
__device__ __forceinline__ void edge_post(const float (&S)[PSI_D], int deg, float (&mp)[PSI_D]) {
    const EdgeMLP& W = edge_mlp<WHICH>();
    const float fdeg = (float)deg;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float t = fdeg * W.b2[o];
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) t = fmaf(W.W2[o][i], S[i], t);
        mp[o] = t;
    }
}

// ---- pre-pass: Q[w][node] = W1j_w · h[node] ------------------------------------------------------------------------
template <int NQ>
__global__ void __launch_bounds__(PSI_NODE_BLOCK) k_layer_pre(int N, const float* __restrict__ h, float* __restrict__ Q, const int* __restrict__ done) {
    if (done != nullptr && *done) return;
    const int node = blockIdx.x * PSI_NODE_BLOCK + threadIdx.x;
    if (node >= N) return;
    float hi[PSI_D], q[PSI_D];
    load_row(h, node, hi);
    edge_q<0>(hi, q);
    store_row(Q, node, q);
    edge_q<1>(hi, q);
    store_row(Q, (int64_t)N + node, q);
    if (NQ > 2) {
        edge_q<2>(hi, q);
        store_row(Q, 2 * (int64_t)N + node, q);
    }
}

// ---- cooperative row gather ---------------------------------------------------------------------------------------
// A trip needs 32 rows of 10 floats (one per lane).  They are fetched as 160 float2 pieces: lane l fetches pieces l + 32p
// (p = 0..4); piece k belongs to row-slot k / 5 (the lane that wants it), column pair k % 5.  Staged at a 12-float row pitch so
// that the read-back is 16-byte aligned and bank-conflict free.
#define PSI_STAGE_PITCH 12
#define PSI_STAGE_FLOATS (32 * PSI_STAGE_PITCH)
struct CoopMap {
    int r[5];      // lane whose row piece p of this lane belongs to: (32p + lane) / 5
    int lane;
};
__device__ __forceinline__ void coop_map(int lane, CoopMap& M) {
    M.lane = lane;
#pragma unroll
    for (int p = 0; p < 5; ++p) M.r[p] = (p * 32 + lane) / 5;
}
// float offset of piece p inside its source row: 2·((32p + lane) − 5r) ; inside the stage: 12r + that = 2r + 2·lane + 64p
__device__ __forceinline__ int coop_src_off(const CoopMap& M, int p) { return 2 * (p * 32 + M.lane - 5 * M.r[p]); }
__device__ __forceinline__ int coop_stage_off(const CoopMap& M, int p) { return 2 * (M.r[p] + M.lane) + 64 * p; }
// rowidx: the row (in units of 10 floats from `src`) this lane wants for the trip, or −1
__device__ __forceinline__ void coop_issue(const float* __restrict__ src, int rowidx, const CoopMap& M, float2 (&v)[5]) {
#pragma unroll
    for (int p = 0; p < 5; ++p) {
        const int rr = __shfl_sync(0xffffffffu, rowidx, M.r[p]);
        v[p] = (rr >= 0) ? __ldg(reinterpret_cast<const float2*>(src + (int64_t)rr * PSI_D + coop_src_off(M, p))) : make_float2(0.f, 0.f);
    }
}
// same through the coherent path (for buffers the compiler must not assume read-only)
__device__ __forceinline__ void coop_issue_rw(const float* src, int rowidx, const CoopMap& M, float2 (&v)[5]) {
#pragma unroll
    for (int p = 0; p < 5; ++p) {
        const int rr = __shfl_sync(0xffffffffu, rowidx, M.r[p]);
        v[p] = (rr >= 0) ? *reinterpret_cast<const float2*>(src + (int64_t)rr * PSI_D + coop_src_off(M, p)) : make_float2(0.f, 0.f);
    }
}
__device__ __forceinline__ void coop_store(float* st, const CoopMap& M, const float2 (&v)[5]) {
#pragma unroll
    for (int p = 0; p < 5; ++p) *reinterpret_cast<float2*>(st + coop_stage_off(M, p)) = v[p];
}
__device__ __forceinline__ void coop_row(const float* st, int lane, float (&q)[PSI_D]) {
    const float4 a = *reinterpret_cast<const float4*>(st + lane * PSI_STAGE_PITCH);
    const float4 b = *reinterpret_cast<const float4*>(st + lane * PSI_STAGE_PITCH + 4);
    const float2 c = *reinterpret_cast<const float2*>(st + lane * PSI_STAGE_PITCH + 8);
    q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w; q[8] = c.x; q[9] = c.y;
}

// Walks the slice column of every lane of the warp over one SELL list.  WARP-UNIFORM: all 32 lanes must call it (the trip count is
// the slice width).  rowidx_of(j) → row index into `src` or −1 when this lane does not want the row; body(rec, row) is called for
// every wanted record in CSR order.
template <class RowIdx, class Body>
__device__ __forceinline__ void walk_list(const SellDev& L, const float* __restrict__ src, int slice, int lane, float* st, const CoopMap& M,
                                          RowIdx&& rowidx_of, Body&& body) {
    const int64_t base = L.slice_off[slice];
    const int width = (int)((L.slice_off[slice + 1] - base) >> 5);
    if (width == 0) return;
    const int4* p = L.recs + base + lane;
    const int4 none = make_int4(-1, 0, 0, 0);
    int4 r0 = __ldg(p);
    int4 r1 = (width > 1) ? __ldg(p + 32) : none;
    int i0 = (r0.x >= 0) ? rowidx_of(r0.x) : -1;
    float2 v[5];
    coop_issue(src, i0, M, v);
    for (int t = 0; t < width; ++t) {
        coop_store(st, M, v);
        __syncwarp();
        const int4 r2 = (t + 2 < width) ? __ldg(p + (int64_t)(t + 2) * 32) : none;
        const int i1 = (r1.x >= 0) ? rowidx_of(r1.x) : -1;
        if (t + 1 < width) coop_issue(src, i1, M, v);          // rows of the next trip fly while this one is consumed
        float q[PSI_D];
        coop_row(st, lane, q);
        __syncwarp();
        if (i0 >= 0) body(r0, q);
        r0 = r1; r1 = r2; i0 = i1;
    }
}

__device__ __forceinline__ float sigmoidf_acc(float s) { return 1.0f / (1.0f + expf(-s)); }

// LayerNorm over the 10 latent channels, eps 1e-5, biased variance (nn.LayerNorm, model.py:270,293).
__device__ __forceinline__ void layer_norm10(const float (&r)[PSI_D], float (&out)[PSI_D], float (&rhat)[PSI_D], float& rstd) {
    float mu = 0.f;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) mu += r[o];
    mu = __fdiv_rn(mu, (float)PSI_D);
    float var = 0.f;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) { float c = r[o] - mu; var = fmaf(c, c, var); }
    var = __fdiv_rn(var, (float)PSI_D);
    rstd = __fdiv_rn(1.0f, __fsqrt_rn(var + 1e-5f));
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        rhat[o] = (r[o] - mu) * rstd;
        out[o] = fmaf(rhat[o], cW.ln_g[o], cW.ln_b[o]);
    }
}

// hidden = relu(up_W1·c + up_b1) with c = [h, to, from, prb(PRB)];  m = up_W2·hidden + up_b2
template <int PRB>
__device__ __forceinline__ void update_mlp(const float (&hi)[PSI_D], const float (&mT)[PSI_D], const float (&mF)[PSI_D],
                                           const float (&prb)[3], float (&m)[PSI_D], uint32_t& hmask, float (&hid)[PSI_D]) {
    hmask = 0;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float t = cW.up_b1[o];
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) t = fmaf(cW.up_W1[o][i], hi[i], t);
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) t = fmaf(cW.up_W1[o][PSI_D + i], mT[i], t);
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) t = fmaf(cW.up_W1[o][2 * PSI_D + i], mF[i], t);
#pragma unroll
        for (int i = 0; i < PRB; ++i) t = fmaf(cW.up_W1[o][3 * PSI_D + i], prb[i], t);
        if (t > 0.f) hmask |= (1u << o);
        hid[o] = fmaxf(t, 0.f);
    }
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float t = cW.up_b2[o];
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) t = fmaf(cW.up_W2[o][i], hid[i], t);
        m[o] = t;
    }
}
template <int PRB>
__device__ __forceinline__ void update_mlp(const float (&hi)[PSI_D], const float (&mT)[PSI_D], const float (&mF)[PSI_D],
                                           const float (&prb)[3], float (&m)[PSI_D], uint32_t& hmask) {
    float hid[PSI_D];
    update_mlp<PRB>(hi, mT, mF, prb, m, hmask, hid);
}

// pre-activation of the gate: s = gate_w·c + gate_b
template <int PRB>
__device__ __forceinline__ float gate_pre(const float (&hi)[PSI_D], const float (&mT)[PSI_D], const float (&mF)[PSI_D],
                                          const float (&prb)[3]) {
    float s = cW.gate_b;
#pragma unroll
    for (int i = 0; i < PSI_D; ++i) s = fmaf(cW.gate_w[i], hi[i], s);
#pragma unroll
    for (int i = 0; i < PSI_D; ++i) s = fmaf(cW.gate_w[PSI_D + i], mT[i], s);
#pragma unroll
    for (int i = 0; i < PSI_D; ++i) s = fmaf(cW.gate_w[2 * PSI_D + i], mF[i], s);
#pragma unroll
    for (int i = 0; i < PRB; ++i) s = fmaf(cW.gate_w[3 * PSI_D + i], prb[i], s);
    return s;
}
template <int PRB>
__device__ __forceinline__ float gate(const float (&hi)[PSI_D], const float (&mT)[PSI_D], const float (&mF)[PSI_D],
                                      const float (&prb)[3]) {
    return sigmoidf_acc(gate_pre<PRB>(hi, mT, mF, prb));
}

// update_neumann: MLP(cat[h, mp_neu, prb(3), normal(2)])  (mixed/psignn/model.py:214,231-232)
__device__ __forceinline__ void neumann_mlp(const float (&hi)[PSI_D], const float (&mN)[PSI_D], const float (&prb)[3],
                                            const float (&nv)[2], float (&m)[PSI_D], uint32_t& hmask, float (&hid)[PSI_D]) {
    hmask = 0;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float t = cW.un_b1[o];
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) t = fmaf(cW.un_W1[o][i], hi[i], t);
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) t = fmaf(cW.un_W1[o][PSI_D + i], mN[i], t);
#pragma unroll
        for (int i = 0; i < 3; ++i) t = fmaf(cW.un_W1[o][2 * PSI_D + i], prb[i], t);
#pragma unroll
        for (int i = 0; i < 2; ++i) t = fmaf(cW.un_W1[o][2 * PSI_D + 3 + i], nv[i], t);
        if (t > 0.f) hmask |= (1u << o);
        hid[o] = fmaxf(t, 0.f);
    }
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float t = cW.un_b2[o];
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) t = fmaf(cW.un_W2[o][i], hid[i], t);
        m[o] = t;
    }
}
__device__ __forceinline__ void neumann_mlp(const float (&hi)[PSI_D], const float (&mN)[PSI_D], const float (&prb)[3],
                                            const float (&nv)[2], float (&m)[PSI_D], uint32_t& hmask) {
    float hid[PSI_D];
    neumann_mlp(hi, mN, prb, nv, m, hmask, hid);
}

template <int PRB>
__device__ __forceinline__ void load_prb(const GraphDev& G, int node, float (&prb)[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) prb[i] = (i < PRB) ? __ldg(G.prb + (int64_t)node * PRB + i) : 0.f;
}

// node class of a lane: 0 interior, 1 Dirichlet, 2 Neumann, 3 not computed (beyond the rows this kernel produces)
template <int KIND>
__device__ __forceinline__ int node_class(const GraphDev& G, int node, int limit) {
    if (node >= limit) return 3;
    if (KIND == KIND_DSS) return 0;
    const uint8_t tg = G.tag[node];
    if (tg & 1) return 1;
    if (KindTraits<KIND>::has_neumann && (tg & 2)) return 2;
    return 0;
}

// Aggregated messages of one node: mT = ΣΦ→ (interior), mF = ΣΦ← (interior) or ΣΦ_neumann (Neumann rows).  WARP-UNIFORM.
template <int KIND>
__device__ __forceinline__ void aggregate(const GraphDev& G, const float* __restrict__ Q, int node, int cls, const float (&hi)[PSI_D],
                                          float* st, const CoopMap& M, float (&mT)[PSI_D], float (&mF)[PSI_D]) {
    constexpr int ATTR = KindTraits<KIND>::ATTR;
    const int lane = threadIdx.x & 31, slice = node >> 5;
    const int N = G.N;
    float P[PSI_D], S[PSI_D];
    int deg = 0;
    // ---- list T: messages Φ→ into interior destinations (neighbour = row index of the entry) ----
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) { S[o] = 0.f; P[o] = 0.f; }
    if (cls == 0) edge_pre<0>(hi, P);
    walk_list(G.T, Q, slice, lane, st, M,
              [&](int j) { return cls == 0 ? j : -1; },
              [&](const int4& rec, const float (&q)[PSI_D]) {
                  float z[PSI_D];
                  edge_z<0, ATTR>(P, q, rec, z);
#pragma unroll
                  for (int o = 0; o < PSI_D; ++o) S[o] += fmaxf(z[o], 0.f);
                  ++deg;
              });
    if (cls == 0) edge_post<0>(S, deg, mT);
    // ---- list F: messages Φ← into interior destinations, Φ_neumann into Neumann destinations ----
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) S[o] = 0.f;
    deg = 0;
    if (cls == 0) edge_pre<1>(hi, P);
    if (KindTraits<KIND>::has_neumann && cls == 2) edge_pre<2>(hi, P);
    walk_list(G.F, Q, slice, lane, st, M,
              [&](int j) { return cls == 0 ? N + j : ((KindTraits<KIND>::has_neumann && cls == 2) ? 2 * N + j : -1); },
              [&](const int4& rec, const float (&q)[PSI_D]) {
                  float z[PSI_D];
                  if (KindTraits<KIND>::has_neumann && cls == 2) edge_z<2, ATTR>(P, q, rec, z);
                  else edge_z<1, ATTR>(P, q, rec, z);
#pragma unroll
                  for (int o = 0; o < PSI_D; ++o) S[o] += fmaxf(z[o], 0.f);
                  ++deg;
              });
    if (cls == 0) edge_post<1>(S, deg, mF);
    if (KindTraits<KIND>::has_neumann && cls == 2) edge_post<2>(S, deg, mF);
}

// GRU-style node update of DSGPS (dirichlet/dsgps/model.py:148-155): H + σ(Z c)·tanh(C·cat[σ(R c)·H, to, from, prb])
template <int PRB>
__device__ __forceinline__ void dsgps_update(const float (&hi)[PSI_D], const float (&mT)[PSI_D], const float (&mF)[PSI_D],
                                             const float (&prb)[3], float (&out)[PSI_D]) {
    float c[33];
#pragma unroll
    for (int i = 0; i < PSI_D; ++i) { c[i] = hi[i]; c[PSI_D + i] = mT[i]; c[2 * PSI_D + i] = mF[i]; }
#pragma unroll
    for (int i = 0; i < 3; ++i) c[30 + i] = prb[i];
    float zk[PSI_D], rk[PSI_D];
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float a = cW.gz_b[o], b = cW.gr_b[o];
#pragma unroll
        for (int i = 0; i < 30 + PRB; ++i) { a = fmaf(cW.gz_W[o][i], c[i], a); b = fmaf(cW.gr_W[o][i], c[i], b); }
        zk[o] = sigmoidf_acc(a);
        rk[o] = sigmoidf_acc(b);
    }
#pragma unroll
    for (int i = 0; i < PSI_D; ++i) c[i] = rk[i] * hi[i];                       // cat[reset*H, to, from, prb]
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float a = cW.gc_b[o];
#pragma unroll
        for (int i = 0; i < 30 + PRB; ++i) a = fmaf(cW.gc_W[o][i], c[i], a);
        out[o] = fmaf(zk[o], tanhf(a), hi[o]);                                  // H + alpha*corr
    }
}

// One application of the layer for one node (after the aggregation).  `hi` is the node's own row of h.
template <int KIND>
__device__ __forceinline__ void node_update(const GraphDev& G, const float* __restrict__ h0, int node, int cls, const float (&hi)[PSI_D],
                                            const float (&mT)[PSI_D], const float (&mF)[PSI_D], float (&out)[PSI_D]) {
    constexpr int PRB = KindTraits<KIND>::PRB;
    if (cls == 1) {                               // Dirichlet clamp: h[dir] = h_initial[dir] (model.py:298)
        load_row(h0, node, out);
        return;
    }
    float prb[3];
    load_prb<PRB>(G, node, prb);
    if (KIND == KIND_DIRICHLET || KIND == KIND_MIXED) {
        float r[PSI_D], m[PSI_D];
        uint32_t hm;
        if (KIND == KIND_MIXED && cls == 2) {     // Neumann rows are overwritten before LayerNorm (mixed model.py:235-237)
            const float nv[2] = {__ldg(G.nrm + 2 * (int64_t)node), __ldg(G.nrm + 2 * (int64_t)node + 1)};
            neumann_mlp(hi, mF, prb, nv, m, hm);
#pragma unroll
            for (int o = 0; o < PSI_D; ++o) r[o] = m[o];
        } else {
            const float alpha = gate<PRB>(hi, mT, mF, prb);
            update_mlp<PRB>(hi, mT, mF, prb, m, hm);
#pragma unroll
            for (int o = 0; o < PSI_D; ++o) r[o] = fmaf(alpha, m[o], hi[o]);
        }
        float rhat[PSI_D], rstd;
        layer_norm10(r, out, rhat, rstd);
    } else if (KIND == KIND_DSS) {
        float m[PSI_D];
        uint32_t hm;
        update_mlp<3>(hi, mT, mF, prb, m, hm);
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) out[o] = fmaf(cW.dss_alpha, m[o], hi[o]);   // H + alpha*Psi (dss model.py:119)
    } else {                                      // DSGPS (dirichlet and mixed)
        if (KIND == KIND_DSGPS_MIXED && cls == 2) {   // H[neumann] = update_neumann[neumann] (mixed/dsgps/model.py:92), no LayerNorm
            const float nv[2] = {__ldg(G.nrm + 2 * (int64_t)node), __ldg(G.nrm + 2 * (int64_t)node + 1)};
            uint32_t hm;
            neumann_mlp(hi, mF, prb, nv, out, hm);
        } else {
            dsgps_update<PRB>(hi, mT, mF, prb, out);
        }
    }
}

// ---- solver epilogue shared by the layer and the VJP operator kernels ---------------------------
// g = op(x) − x ; δg = g − g_old ; partial sums of ‖g‖² and ‖g + x‖²  (solver.py:123,131,162-163)
struct SolverEpi {
    float* g;            // [numel] in/out
    float* dg;           // [numel] out
    float* norm_part;    // [2, gridDim.x] out
    const int* done;     // device flag: skip all work once the solve has finished
};

__device__ __forceinline__ void solver_epilogue(const SolverEpi& E, int node, bool valid, const float (&xi)[PSI_D],
                                                const float (&fx)[PSI_D], float* smem) {
    float acc[2] = {0.f, 0.f};
    if (valid) {
        float gold[PSI_D], gn[PSI_D], dgv[PSI_D];
        load_row_rw(E.g, node, gold);
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) {
            gn[o] = fx[o] - xi[o];
            dgv[o] = gn[o] - gold[o];
            const float fr = gn[o] + xi[o];
            acc[0] = fmaf(gn[o], gn[o], acc[0]);
            acc[1] = fmaf(fr, fr, acc[1]);
        }
        store_row(E.g, node, gn);
        store_row(E.dg, node, dgv);
    }
    block_sum<2, PSI_NODE_BLOCK / 32>(acc, smem);
    if (threadIdx.x == 0) {
        E.norm_part[blockIdx.x] = acc[0];
        E.norm_part[gridDim.x + blockIdx.x] = acc[1];
    }
}

template <int KIND, bool EPI>
__global__ void __launch_bounds__(PSI_NODE_BLOCK, PSI_OP_MIN_CTAS)
k_layer_forward(GraphDev G, const float* __restrict__ h, const float* __restrict__ h0, const float* __restrict__ Q, float* __restrict__ out,
                SolverEpi E) {
    __shared__ __align__(16) float stage[(PSI_NODE_BLOCK / 32) * PSI_STAGE_FLOATS];
    __shared__ float smem[2 * PSI_NODE_BLOCK / 32];
    if (EPI && *E.done) return;
    const int node = blockIdx.x * PSI_NODE_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = node < G.n_compute;
    float hi[PSI_D], fx[PSI_D];
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) { hi[o] = 0.f; fx[o] = 0.f; }
    // a warp whose slice lies entirely beyond the produced rows has nothing to gather (warp-uniform condition)
    if ((node & ~31) < G.n_compute) {
        const int cls = node_class<KIND>(G, node, G.n_compute);
        if (valid) load_row(h, node, hi);
        CoopMap M;
        coop_map(lane, M);
        float mT[PSI_D], mF[PSI_D];
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) { mT[o] = 0.f; mF[o] = 0.f; }
        aggregate<KIND>(G, Q, node, cls, hi, stage + (threadIdx.x >> 5) * PSI_STAGE_FLOATS, M, mT, mF);
        if (valid) {
            node_update<KIND>(G, h0, node, cls, hi, mT, mF, fx);
            if (!EPI || out != nullptr) store_row(out, node, fx);   // Picard keeps f(x) itself
        }
    }
    if (EPI) solver_epilogue(E, node, valid, hi, fx, smem);
}

// ---- encoder / decoder (model.py:370-389) --------------------------------------------------------
__global__ void __launch_bounds__(PSI_NODE_BLOCK) k_encode(int N, const float* __restrict__ x, float* __restrict__ h) {
    const int node = blockIdx.x * PSI_NODE_BLOCK + threadIdx.x;
    if (node >= N) return;
    const float xv = __ldg(x + node);
    float hid[PSI_D], o_[PSI_D];
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) hid[o] = fmaxf(fmaf(cW.enc_W1[o], xv, cW.enc_b1[o]), 0.f);
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float t = cW.enc_b2[o];
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) t = fmaf(cW.enc_W2[o][i], hid[i], t);
        o_[o] = t;
    }
    store_row(h, node, o_);
}

__global__ void __launch_bounds__(PSI_NODE_BLOCK) k_decode(int N, const float* __restrict__ h, float* __restrict__ u) {
    const int node = blockIdx.x * PSI_NODE_BLOCK + threadIdx.x;
    if (node >= N) return;
    float hi[PSI_D];
    load_row(h, node, hi);
    float acc = cW.dec_b2;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float t = cW.dec_b1[o];
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) t = fmaf(cW.dec_W1[o][i], hi[i], t);
        acc = fmaf(cW.dec_W2[o], fmaxf(t, 0.f), acc);
    }
    u[node] = acc;
}

// ---- residual SpMV  r = A u − y  and  Σ r²  (model.py:157-167) ----------------------------------
__global__ void __launch_bounds__(PSI_NODE_BLOCK)
k_residual(GraphDev G, const float* __restrict__ u, const float* __restrict__ y, float* __restrict__ r, float* __restrict__ part) {
    __shared__ float smem[PSI_NODE_BLOCK / 32];
    const int node = blockIdx.x * PSI_NODE_BLOCK + threadIdx.x;
    float acc[1] = {0.f};
    if (node < G.N) {
        const SellDev& L = G.Ar;
        const int64_t base = L.slice_off[node >> 5];
        const int width = (int)((L.slice_off[(node >> 5) + 1] - base) >> 5);
        const int2* p = L.recs2 + base + (node & 31);
        float s = 0.f;
        for (int t = 0; t < width; ++t) {
            const int2 rec = __ldg(p + (int64_t)t * 32);
            if (rec.x >= 0) s = fmaf(__int_as_float(rec.y), __ldg(u + rec.x), s);
        }
        const float res = s - __ldg(y + node);
        if (r != nullptr) r[node] = res;
        acc[0] = res * res;
    }
    block_sum<1, PSI_NODE_BLOCK / 32>(acc, smem);
    if (threadIdx.x == 0) part[blockIdx.x] = acc[0];
}

// out = Aᵀ v  (column-grouped list)
__global__ void __launch_bounds__(PSI_NODE_BLOCK) k_spmv_t(GraphDev G, const float* __restrict__ v, float* __restrict__ out) {
    const int node = blockIdx.x * PSI_NODE_BLOCK + threadIdx.x;
    if (node >= G.N) return;
    const SellDev& L = G.Ac;
    const int64_t base = L.slice_off[node >> 5];
    const int width = (int)((L.slice_off[(node >> 5) + 1] - base) >> 5);
    const int2* p = L.recs2 + base + (node & 31);
    float s = 0.f;
    for (int t = 0; t < width; ++t) {
        const int2 rec = __ldg(p + (int64_t)t * 32);
        if (rec.x >= 0) s = fmaf(__int_as_float(rec.y), __ldg(v + rec.x), s);
    }
    out[node] = s;
}

// deterministic final reduction of per-block partial sums: out[0] = scale * Σ part
__global__ void k_reduce_partials(int n, const float* __restrict__ part, float scale, float* __restrict__ out) {
    __shared__ double sm[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)part[i];
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
        out[0] = (float)(t * (double)scale);
    }
}
