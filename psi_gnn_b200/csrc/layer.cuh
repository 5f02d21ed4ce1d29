// layer.cuh — the fused message-passing layer f_theta (one kernel per application).
//
// One thread owns one destination node (one warp = one 32-node slice of the SELL lists):
//   1. coalesced 16-byte edge records {j, a0, a1, a2}; 8-byte vectorised gather of the neighbour row h_j
//   2. edge MLP Phi: first layer split by input block (W1i·h_i + b1 hoisted per destination),
//      ReLU, summed per destination in CSR order (deterministic, no atomics); the second edge layer
//      is applied once to the sum:  Σ_e (W2·relu(z_e) + b2) = W2·Σ_e relu(z_e) + deg·b2
//   3. node update Psi (gate · MLP), LayerNorm, boundary masks (Dirichlet clamp, Neumann overwrite)
//   4. optional solver epilogue: g = f(x) − x, δg, and the two stopping norms
// All weights are constant-bank FFMA operands (weights.cuh).
//
// Reference semantics: dirichlet/psignn/model.py:279-300 (+ :334-368), mixed/psignn/model.py:216-245,
// dirichlet/dss/model.py:113-121, dirichlet/dsgps/model.py:143-163.
#pragma once
#include "common.cuh"
#include "weights.cuh"
#include "graph.cuh"

enum { KIND_DIRICHLET = 0, KIND_MIXED = 1, KIND_DSS = 2, KIND_DSGPS = 3 };

// z-chain of the first edge layer.  The summation order b1 → W1i·h_dst → W1j·h_src → W1a·a is fixed
// here so that every kernel that needs z (forward, VJP-prepare own and cross masks) rounds identically.
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
template <int WHICH, int ATTR>
__device__ __forceinline__ void edge_z(const float (&P)[PSI_D], const float (&hs)[PSI_D], const int4& rec, float (&z)[PSI_D]) {
    const EdgeMLP& W = edge_mlp<WHICH>();
    const float a0 = __int_as_float(rec.y), a1 = __int_as_float(rec.z), a2 = __int_as_float(rec.w);
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float t = P[o];
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) t = fmaf(W.W1j[o][i], hs[i], t);
        t = fmaf(W.W1a[o][0], a0, t);
        if (ATTR > 1) t = fmaf(W.W1a[o][1], a1, t);
        if (ATTR > 2) t = fmaf(W.W1a[o][2], a2, t);
        z[o] = t;
    }
}

// Σ_e relu(z_e) over the node's slice column, then the second edge layer.
template <int WHICH, int ATTR>
__device__ __forceinline__ void edge_aggregate(const SellDev& L, const float* __restrict__ h, int node,
                                               const float (&hi)[PSI_D], float (&mp)[PSI_D]) {
    const EdgeMLP& W = edge_mlp<WHICH>();
    float P[PSI_D], S[PSI_D];
    edge_pre<WHICH>(hi, P);
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) S[o] = 0.f;
    const int64_t base = L.slice_off[node >> 5];
    const int width = (int)((L.slice_off[(node >> 5) + 1] - base) >> 5);
    const int4* p = L.recs + base + (node & 31);
    int deg = 0;
    // two edges per trip: both records, then both neighbour rows, are requested before the FMAs of either edge start, which
    // halves the number of dependent memory round trips of the per-destination loop (the kernel is latency-, not FMA-bound)
    for (int t0 = 0; t0 < width; t0 += 2) {
        int4 rec[2];
        rec[0] = __ldg(p + (int64_t)t0 * 32);
        rec[1] = (t0 + 1 < width) ? __ldg(p + (int64_t)(t0 + 1) * 32) : make_int4(-1, 0, 0, 0);
        float hj[2][PSI_D];
#pragma unroll
        for (int q = 0; q < 2; ++q)
            if (rec[q].x >= 0) load_row(h, rec[q].x, hj[q]);
#pragma unroll
        for (int q = 0; q < 2; ++q)
            if (rec[q].x >= 0) {
                float z[PSI_D];
                edge_z<WHICH, ATTR>(P, hj[q], rec[q], z);
#pragma unroll
                for (int o = 0; o < PSI_D; ++o) S[o] += fmaxf(z[o], 0.f);
                ++deg;
            }
    }
    const float fdeg = (float)deg;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float t = fdeg * W.b2[o];
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) t = fmaf(W.W2[o][i], S[i], t);
        mp[o] = t;
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
                                           const float (&prb)[3], float (&m)[PSI_D], uint32_t& hmask) {
    float hid[PSI_D];
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
__device__ __forceinline__ float gate(const float (&hi)[PSI_D], const float (&mT)[PSI_D], const float (&mF)[PSI_D],
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
    return sigmoidf_acc(s);
}

// update_neumann: MLP(cat[h, mp_neu, prb(3), normal(2)])  (mixed/psignn/model.py:214,231-232)
__device__ __forceinline__ void neumann_mlp(const float (&hi)[PSI_D], const float (&mN)[PSI_D], const float (&prb)[3],
                                            const float (&nv)[2], float (&m)[PSI_D], uint32_t& hmask) {
    float hid[PSI_D];
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

template <int PRB>
__device__ __forceinline__ void load_prb(const GraphDev& G, int node, float (&prb)[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) prb[i] = (i < PRB) ? __ldg(G.prb + (int64_t)node * PRB + i) : 0.f;
}

// One application of the layer for one node.  `hi` is the node's own row of h.
template <int KIND>
__device__ __forceinline__ void node_forward(const GraphDev& G, const float* __restrict__ h, const float* __restrict__ h0,
                                             int node, const float (&hi)[PSI_D], float (&out)[PSI_D]) {
    const uint8_t tg = G.tag[node];
    if (KIND != KIND_DSS && (tg & 1)) {           // Dirichlet clamp: h[dir] = h_initial[dir] (model.py:298)
        load_row(h0, node, out);
        return;
    }
    if (KIND == KIND_DIRICHLET || KIND == KIND_MIXED) {
        constexpr int PRB = (KIND == KIND_MIXED) ? 3 : 2;
        float prb[3];
        load_prb<PRB>(G, node, prb);
        float r[PSI_D];
        if (KIND == KIND_MIXED && (tg & 2)) {     // Neumann rows are overwritten before LayerNorm (mixed model.py:235-237)
            float mN[PSI_D], m[PSI_D];
            edge_aggregate<2, 3>(G.F, h, node, hi, mN);
            const float nv[2] = {__ldg(G.nrm + 2 * (int64_t)node), __ldg(G.nrm + 2 * (int64_t)node + 1)};
            uint32_t hm;
            neumann_mlp(hi, mN, prb, nv, m, hm);
#pragma unroll
            for (int o = 0; o < PSI_D; ++o) r[o] = m[o];
        } else {
            float mT[PSI_D], mF[PSI_D], m[PSI_D];
            edge_aggregate<0, 3>(G.T, h, node, hi, mT);
            edge_aggregate<1, 3>(G.F, h, node, hi, mF);
            const float alpha = gate<PRB>(hi, mT, mF, prb);
            uint32_t hm;
            update_mlp<PRB>(hi, mT, mF, prb, m, hm);
#pragma unroll
            for (int o = 0; o < PSI_D; ++o) r[o] = fmaf(alpha, m[o], hi[o]);
        }
        float rhat[PSI_D], rstd;
        layer_norm10(r, out, rhat, rstd);
    } else if (KIND == KIND_DSS) {
        float prb[3], mT[PSI_D], mF[PSI_D], m[PSI_D];
        load_prb<3>(G, node, prb);
        edge_aggregate<0, 1>(G.T, h, node, hi, mT);
        edge_aggregate<1, 1>(G.F, h, node, hi, mF);
        uint32_t hm;
        update_mlp<3>(hi, mT, mF, prb, m, hm);
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) out[o] = fmaf(cW.dss_alpha, m[o], hi[o]);   // H + alpha*Psi (dss model.py:119)
    } else {  // KIND_DSGPS
        float prb[3], mT[PSI_D], mF[PSI_D];
        load_prb<2>(G, node, prb);
        edge_aggregate<0, 3>(G.T, h, node, hi, mT);
        edge_aggregate<1, 3>(G.F, h, node, hi, mF);
        float c[32];
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) { c[i] = hi[i]; c[PSI_D + i] = mT[i]; c[2 * PSI_D + i] = mF[i]; }
        c[30] = prb[0]; c[31] = prb[1];
        float zk[PSI_D], rk[PSI_D];
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) {
            float a = cW.gz_b[o], b = cW.gr_b[o];
#pragma unroll
            for (int i = 0; i < 32; ++i) { a = fmaf(cW.gz_W[o][i], c[i], a); b = fmaf(cW.gr_W[o][i], c[i], b); }
            zk[o] = sigmoidf_acc(a);
            rk[o] = sigmoidf_acc(b);
        }
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) c[i] = rk[i] * hi[i];                       // cat[reset*H, to, from, prb]
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) {
            float a = cW.gc_b[o];
#pragma unroll
            for (int i = 0; i < 32; ++i) a = fmaf(cW.gc_W[o][i], c[i], a);
            out[o] = fmaf(zk[o], tanhf(a), hi[o]);                                  // H + alpha*corr (dsgps model.py:152-155)
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
__global__ void __launch_bounds__(PSI_NODE_BLOCK)
k_layer_forward(GraphDev G, const float* __restrict__ h, const float* __restrict__ h0, float* __restrict__ out, SolverEpi E) {
    __shared__ float smem[2 * PSI_NODE_BLOCK / 32];
    if (EPI && *E.done) return;
    const int node = blockIdx.x * PSI_NODE_BLOCK + threadIdx.x;
    const bool valid = node < G.n_compute;
    float hi[PSI_D], fx[PSI_D];
    if (valid) {
        load_row(h, node, hi);
        node_forward<KIND>(G, h, h0, node, hi, fx);
        if (!EPI || out != nullptr) store_row(out, node, fx);   // Picard keeps f(x) itself
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
