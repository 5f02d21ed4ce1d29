// pgrad.cuh — parameter gradient of one application of f_theta at a frozen point:  θ̄ = (∂f/∂θ at H*)ᵀ ȳ.
//
// Replaces the autograd walk from the single differentiable re-application f(H*) of a training step to the parameters
// (dirichlet/psignn/model.py:204-205 + loss.backward(): addmm / index_add / layer_norm backward kernels, whose index_add and
// index_select backward use float atomics on CUDA).  Here every parameter gradient is a sum over nodes of a product of two per-node
// quantities, formed in a fixed order: deterministic, no atomics.
//
//   θ̄[p] = Σ_nodes  rec[node][ty(p)] · rec[node][tx(p)]
//
// where `rec` is a per-node record of forward intermediates (recomputed from H*) and cotangents (the algebra of vjp.cuh phase A),
// and (ty, tx) come from a table built by psi_gnn_b200/weights.py from the LayerWeights layout.  Scatter → gather once more: the
// first-layer source block W1j of an edge MLP needs Σ_e z̄_e ⊗ h_src(e); regrouped by SOURCE node it is acc_src ⊗ h_src with
// acc = Σ_{edges leaving the node} mask_e ⊙ S̄_dst — exactly what VJP phase B gathers (it stores them when asked to).
//
// Record layout (floats); psi_pgrad_layout() reports it so that the Python table builder can check itself against it.
#pragma once
#include "vjp.cuh"

enum {
    PG_ONE = 0,        // 1.0 (bias columns)
    PG_DEG = 1,        // deg_T, deg_F, deg_N
    PG_C = 4,          // c = [h(10), mT(10), mF(10), prb(3)]                 (interior rows)
    PG_CN = 37,        // cN = [h(10), mN(10), prb(3), normal(2)]             (Neumann rows)
    PG_YB = 62,        // ȳ
    PG_RHAT = 72,      // r̂
    PG_MB = 82,        // m̄ = α r̄                                            (interior)
    PG_HID = 92,       // hidden of the update MLP
    PG_TB = 102,       // t̄ = hid̄ ⊙ [t > 0]
    PG_SB = 112,       // s̄ (gate pre-activation cotangent)
    PG_EDGE = 113,     // 3 blocks (to, from, neumann) of 70: m̄X(10) S_X(10) zs_X = S̄X⊙cnt_X (10) S̄X(10) Aat_X[10][3]
    PG_ACC = 323,      // accT, accF, accN (30): gathered at the node as SOURCE (phase B)
    PG_MBN = 353,      // Neumann: m̄ = r̄
    PG_HIDN = 363,
    PG_TBN = 373,
    PG_REC = 383,
    PG_PITCH = 385     // odd pitch: rows of different nodes fall into different banks
};
#define PG_NODES 64
#define PG_MAX_PER_THREAD 36             // parameters per thread: table entries ≤ 64·36 = 2304 (mixed PSI-GNN layer: 2 134)

// own-direction aggregate with everything the parameter gradient needs: S = Σ relu(z), activity counts, mask-weighted attribute
// sums Aat[o][c] = Σ_e [z_e[o] > 0]·a_e[c], degree, and the aggregated message mp
template <int OWN>
__device__ __forceinline__ void pgrad_list(const SellDev& L, const float* __restrict__ h, int node, const float (&hi)[PSI_D], float* rec_edge,
                                           float& deg_out, float (&mp)[PSI_D]) {
    float P[PSI_D], S[PSI_D], cnt[PSI_D], A[PSI_D][3];
    edge_pre<OWN>(hi, P);
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) { S[o] = 0.f; cnt[o] = 0.f; A[o][0] = 0.f; A[o][1] = 0.f; A[o][2] = 0.f; }
    const int64_t base = L.slice_off[node >> 5];
    const int width = (int)((L.slice_off[(node >> 5) + 1] - base) >> 5);
    const int4* p = L.recs + base + (node & 31);
    int deg = 0;
    for (int t = 0; t < width; ++t) {
        const int4 r = __ldg(p + (int64_t)t * 32);
        if (r.x < 0) continue;
        float hj[PSI_D], Qv[PSI_D], z[PSI_D];
        load_row(h, r.x, hj);
        edge_q<OWN>(hj, Qv);
        edge_z<OWN, 3>(P, Qv, r, z);
        const float a0 = __int_as_float(r.y), a1 = __int_as_float(r.z), a2 = __int_as_float(r.w);
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) {
            const bool on = z[o] > 0.f;
            S[o] += fmaxf(z[o], 0.f);
            cnt[o] += on ? 1.f : 0.f;
            A[o][0] += on ? a0 : 0.f; A[o][1] += on ? a1 : 0.f; A[o][2] += on ? a2 : 0.f;
        }
        ++deg;
    }
    edge_post<OWN>(S, deg, mp);
    deg_out = (float)deg;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        rec_edge[10 + o] = S[o];
        rec_edge[20 + o] = cnt[o];            // overwritten by zs = S̄ ⊙ cnt once S̄ is known
        rec_edge[40 + 3 * o] = A[o][0]; rec_edge[41 + 3 * o] = A[o][1]; rec_edge[42 + 3 * o] = A[o][2];
    }
}

// fills rec[0 .. PG_REC) of one node
template <int KIND>
__device__ __forceinline__ void pgrad_node(const GraphDev& G, const VjpCacheDev& C, const float* __restrict__ h, const float* __restrict__ y,
                                           const float* __restrict__ acc, int node, float* rec) {
    constexpr int PRB = (KIND == KIND_MIXED) ? 3 : 2;
    for (int i = 0; i < PG_REC; ++i) rec[i] = 0.f;
    if (node >= G.n_compute) return;
    rec[PG_ONE] = 1.f;
    float hi[PSI_D];
    load_row(h, node, hi);
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) { rec[PG_C + o] = hi[o]; rec[PG_CN + o] = hi[o]; }
    for (int i = 0; i < 30; ++i) rec[PG_ACC + i] = acc[(int64_t)node * 30 + i];
    const uint8_t tg = G.tag[node];
    if (tg & 1) return;                                   // Dirichlet row: f = h0, no parameter in it (it still SENDS messages: acc ⊗ h)
    const bool neu = (KIND == KIND_MIXED) && (tg & 2);
    float yi[PSI_D], rhat[PSI_D], rbar[PSI_D];
    load_row(y, node, yi);
    load_row(C.rhat, node, rhat);
    const float rstd = C.rstd[node];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        rec[PG_YB + o] = yi[o]; rec[PG_RHAT + o] = rhat[o];
        rbar[o] = cW.ln_g[o] * yi[o];
        s1 += rbar[o];
        s2 = fmaf(rbar[o], rhat[o], s2);
    }
    s1 = __fdiv_rn(s1, (float)PSI_D);
    s2 = __fdiv_rn(s2, (float)PSI_D);
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) rbar[o] = rstd * (rbar[o] - s1 - rhat[o] * s2);
    float prb[3];
    load_prb<PRB>(G, node, prb);
    if (neu) {
        float mN[PSI_D], degN;
        float* eN = rec + PG_EDGE + 140;
        pgrad_list<2>(G.F, h, node, hi, eN, degN, mN);
        rec[PG_DEG + 2] = degN;
        const float nv[2] = {__ldg(G.nrm + 2 * (int64_t)node), __ldg(G.nrm + 2 * (int64_t)node + 1)};
        float m[PSI_D], hid[PSI_D];
        uint32_t hm;
        neumann_mlp(hi, mN, prb, nv, m, hm, hid);
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) rec[PG_CN + 10 + o] = mN[o];
        rec[PG_CN + 20] = prb[0]; rec[PG_CN + 21] = prb[1]; rec[PG_CN + 22] = prb[2]; rec[PG_CN + 23] = nv[0]; rec[PG_CN + 24] = nv[1];
        float tb[PSI_D], mpb[PSI_D];
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) {
            rec[PG_MBN + o] = rbar[o]; rec[PG_HIDN + o] = hid[o];
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < PSI_D; ++q) t = fmaf(cW.un_W2[q][o], rbar[q], t);
            tb[o] = ((hm >> o) & 1u) ? t : 0.f;
            rec[PG_TBN + o] = tb[o];
        }
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) {
            float b = 0.f;
#pragma unroll
            for (int o = 0; o < PSI_D; ++o) b = fmaf(cW.un_W1[o][PSI_D + i], tb[o], b);
            mpb[i] = b;
        }
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) {
            float t = 0.f;
#pragma unroll
            for (int o = 0; o < PSI_D; ++o) t = fmaf(cW.neu.W2[o][i], mpb[o], t);
            eN[i] = mpb[i];
            eN[30 + i] = t;
            eN[20 + i] = t * eN[20 + i];
        }
        return;
    }
    float mT[PSI_D], mF[PSI_D], degT, degF;
    float* eT = rec + PG_EDGE;
    float* eF = rec + PG_EDGE + 70;
    pgrad_list<0>(G.T, h, node, hi, eT, degT, mT);
    pgrad_list<1>(G.F, h, node, hi, eF, degF, mF);
    rec[PG_DEG + 0] = degT; rec[PG_DEG + 1] = degF;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) { rec[PG_C + 10 + o] = mT[o]; rec[PG_C + 20 + o] = mF[o]; }
    rec[PG_C + 30] = prb[0]; rec[PG_C + 31] = prb[1]; rec[PG_C + 32] = (PRB > 2) ? prb[2] : 0.f;
    const float alpha = gate<PRB>(hi, mT, mF, prb);
    float m[PSI_D], hid[PSI_D];
    uint32_t hm;
    update_mlp<PRB>(hi, mT, mF, prb, m, hm, hid);
    float abar = 0.f;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) abar = fmaf(rbar[o], m[o], abar);
    const float sbar = abar * alpha * (1.0f - alpha);
    rec[PG_SB] = sbar;
    float tb[PSI_D];
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        rec[PG_MB + o] = alpha * rbar[o];
        rec[PG_HID + o] = hid[o];
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < PSI_D; ++q) t = fmaf(cW.up_W2[q][o], alpha * rbar[q], t);
        tb[o] = ((hm >> o) & 1u) ? t : 0.f;
        rec[PG_TB + o] = tb[o];
    }
    float mTb[PSI_D], mFb[PSI_D];
#pragma unroll
    for (int i = 0; i < PSI_D; ++i) {
        float b = sbar * cW.gate_w[PSI_D + i], c = sbar * cW.gate_w[2 * PSI_D + i];
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) {
            b = fmaf(cW.up_W1[o][PSI_D + i], tb[o], b);
            c = fmaf(cW.up_W1[o][2 * PSI_D + i], tb[o], c);
        }
        mTb[i] = b; mFb[i] = c;
    }
#pragma unroll
    for (int i = 0; i < PSI_D; ++i) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) {
            a = fmaf(cW.to.W2[o][i], mTb[o], a);
            b = fmaf(cW.from.W2[o][i], mFb[o], b);
        }
        eT[i] = mTb[i]; eT[30 + i] = a; eT[20 + i] = a * eT[20 + i];
        eF[i] = mFb[i]; eF[30 + i] = b; eF[20 + i] = b * eF[20 + i];
    }
}

// Every CTA walks batches of PG_NODES nodes: the node records go to shared memory, then thread t accumulates the table entries
// t, t + 64, … over the batch.  One row of partial sums per CTA (fixed assignment of batches to CTAs: deterministic).
template <int KIND>
__global__ void __launch_bounds__(PG_NODES)
k_pgrad(GraphDev G, VjpCacheDev C, const float* __restrict__ h, const float* __restrict__ y, const float* __restrict__ acc,
        const int* __restrict__ tab_y, const int* __restrict__ tab_x, int n_tab, float* __restrict__ partial, int num_batches) {
    extern __shared__ float rec[];                         // [PG_NODES][PG_PITCH]
    float a[PG_MAX_PER_THREAD];
    int ty[PG_MAX_PER_THREAD], tx[PG_MAX_PER_THREAD];
#pragma unroll
    for (int k = 0; k < PG_MAX_PER_THREAD; ++k) {
        a[k] = 0.f;
        const int p = threadIdx.x + PG_NODES * k;
        ty[k] = (p < n_tab) ? tab_y[p] : 0;
        tx[k] = (p < n_tab) ? tab_x[p] : 0;
    }
    for (int batch = blockIdx.x; batch < num_batches; batch += gridDim.x) {
        pgrad_node<KIND>(G, C, h, y, acc, batch * PG_NODES + threadIdx.x, rec + threadIdx.x * PG_PITCH);
        __syncthreads();
        for (int nd = 0; nd < PG_NODES; ++nd) {
            const float* r = rec + nd * PG_PITCH;
#pragma unroll
            for (int k = 0; k < PG_MAX_PER_THREAD; ++k) a[k] = fmaf(r[ty[k]], r[tx[k]], a[k]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < PG_MAX_PER_THREAD; ++k) {
        const int p = threadIdx.x + PG_NODES * k;
        if (p < n_tab) partial[(int64_t)blockIdx.x * n_tab + p] = a[k];
    }
}

// θ̄[dst[p]] = Σ_CTA partial[CTA][p]  (fp64, CTA order)
__global__ void __launch_bounds__(128)
k_pgrad_reduce(const float* __restrict__ partial, int n_rows, int n_tab, const int* __restrict__ tab_dst, float* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_tab) return;
    double s = 0.0;
    for (int r = 0; r < n_rows; ++r) s += (double)partial[(int64_t)r * n_tab + p];
    out[tab_dst[p]] = (float)s;
}

// =====================================================================================================================================
// Tangent of the parameter gradient along a direction ḣ of the frozen point:   θ̄'[p] = d/dε θ̄[p](H* + ε ḣ; ȳ)  at ε = 0.
//
// This is the double backward of the Hutchinson regulariser (dirichlet/psignn/model.py:207, :416-435 — autograd.grad(f0, z0, v,
// create_graph=True) followed by loss.backward()):   ∇θ ‖Jᵀv‖² = 2 (Jᵀv)ᵀ ∂θ(Jᵀv) = 2 d/dε ∂θ[vᵀ f_θ(H* + ε w̄)]  with w̄ = Jᵀv held
// fixed, i.e. 2 × the tangent of psi_param_grad(ȳ = v) along ḣ = w̄ (forward-over-reverse).  θ̄[p] is bilinear in the node record,
// so θ̄'[p] = Σ_nodes rec'[ty]·rec[tx] + rec[ty]·rec'[tx]: the record comes from pgrad_node (bit-identical to the primal gradient), its
// tangent rec' from the forward-mode formulas below.  ReLU masks and activity counts are piecewise constant (zero tangent), ȳ does
// not depend on ε.  Two passes, because the tangent of acc (the cotangents S̄ of a node's DESTINATIONS) needs every node's S̄' first:
//   pass 0 stores S̄' per node, k_vjp_phase_b gathers it over the cached masks (same kernel, other input), pass 1 accumulates.
// =====================================================================================================================================

// own-direction walk: S' = Σ_e [z_e > 0] ⊙ (W1i·ḣ_i + W1j·ḣ_j), activity counts (recomputed: pgrad_node overwrote them with zs)
template <int OWN>
__device__ __forceinline__ void pgrad_list_tan(const SellDev& L, const float* __restrict__ h, const float* __restrict__ hdot, int node,
                                               const float (&hi)[PSI_D], const float (&hdi)[PSI_D], float (&St)[PSI_D], float (&cnt)[PSI_D]) {
    const EdgeMLP& Wm = edge_mlp<OWN>();
    float P[PSI_D], Pt[PSI_D];
    edge_pre<OWN>(hi, P);
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) t = fmaf(Wm.W1i[o][i], hdi[i], t);
        Pt[o] = t; St[o] = 0.f; cnt[o] = 0.f;
    }
    const int64_t base = L.slice_off[node >> 5];
    const int width = (int)((L.slice_off[(node >> 5) + 1] - base) >> 5);
    const int4* p = L.recs + base + (node & 31);
    for (int t = 0; t < width; ++t) {
        const int4 r = __ldg(p + (int64_t)t * 32);
        if (r.x < 0) continue;
        float hj[PSI_D], hdj[PSI_D], Qv[PSI_D], z[PSI_D];
        load_row(h, r.x, hj);
        load_row(hdot, r.x, hdj);
        edge_q<OWN>(hj, Qv);
        edge_z<OWN, 3>(P, Qv, r, z);
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) {
            float zt = Pt[o];
#pragma unroll
            for (int i = 0; i < PSI_D; ++i) zt = fmaf(Wm.W1j[o][i], hdj[i], zt);
            const bool on = z[o] > 0.f;
            St[o] += on ? zt : 0.f;
            cnt[o] += on ? 1.f : 0.f;
        }
    }
}

// LayerNorm forward + backward tangents shared by interior and Neumann rows.  In: rhat, rstd (primal), r' (tangent of the pre-norm
// row), r̄ (primal cotangent of the pre-norm row), ȳ.  Out: rhat', r̄'.
__device__ __forceinline__ void ln_tan(const float (&rhat)[PSI_D], float rstd, const float (&rt)[PSI_D], const float (&rbar)[PSI_D],
                                       const float (&yi)[PSI_D], float (&rhat_t)[PSI_D], float (&rbar_t)[PSI_D]) {
    float mu_t = 0.f;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) mu_t += rt[o];
    mu_t = __fdiv_rn(mu_t, (float)PSI_D);
    float q = 0.f;                                     // mean(rhat ⊙ c'),  c' = r' − mean(r')
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) q = fmaf(rhat[o], rt[o] - mu_t, q);
    q = __fdiv_rn(q, (float)PSI_D);
    float s2 = 0.f, s2t = 0.f;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        rhat_t[o] = rstd * ((rt[o] - mu_t) - rhat[o] * q);
        const float gy = cW.ln_g[o] * yi[o];
        s2 = fmaf(gy, rhat[o], s2);
        s2t = fmaf(gy, rhat_t[o], s2t);
    }
    s2 = __fdiv_rn(s2, (float)PSI_D);
    s2t = __fdiv_rn(s2t, (float)PSI_D);
    const float lr = -rstd * q;                        // rstd'/rstd
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) rbar_t[o] = lr * rbar[o] - rstd * (rhat_t[o] * s2 + rhat[o] * s2t);
}

// fills rt[0 .. PG_REC) = tangent of rec (rec already filled by pgrad_node for the same node)
template <int KIND>
__device__ __noinline__ void pgrad_node_tan(const GraphDev& G, const VjpCacheDev& C, const float* __restrict__ h, const float* __restrict__ hdot,
                                               const float* __restrict__ acc_t, int node, const float* rec, float* rt) {
    constexpr int PRB = (KIND == KIND_MIXED) ? 3 : 2;
    for (int i = 0; i < PG_REC; ++i) rt[i] = 0.f;
    if (node >= G.n_compute) return;
    float hi[PSI_D], hdi[PSI_D];
    load_row(h, node, hi);
    load_row(hdot, node, hdi);
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) { rt[PG_C + o] = hdi[o]; rt[PG_CN + o] = hdi[o]; }
    if (acc_t != nullptr)
        for (int i = 0; i < 30; ++i) rt[PG_ACC + i] = acc_t[(int64_t)node * 30 + i];
    const uint8_t tg = G.tag[node];
    if (tg & 1) return;
    const bool neu = (KIND == KIND_MIXED) && (tg & 2);
    float yi[PSI_D], rhat[PSI_D], rbar[PSI_D];
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) { yi[o] = rec[PG_YB + o]; rhat[o] = rec[PG_RHAT + o]; }
    const float rstd = C.rstd[node];
    {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) {
            rbar[o] = cW.ln_g[o] * yi[o];
            s1 += rbar[o];
            s2 = fmaf(rbar[o], rhat[o], s2);
        }
        s1 = __fdiv_rn(s1, (float)PSI_D);
        s2 = __fdiv_rn(s2, (float)PSI_D);
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) rbar[o] = rstd * (rbar[o] - s1 - rhat[o] * s2);
    }
    float prb[3];
    load_prb<PRB>(G, node, prb);
    if (neu) {
        float SNt[PSI_D], cntN[PSI_D], mNt[PSI_D], mN[PSI_D];
        pgrad_list_tan<2>(G.F, h, hdot, node, hi, hdi, SNt, cntN);
        float* eNt = rt + PG_EDGE + 140;
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) {
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < PSI_D; ++i) t = fmaf(cW.neu.W2[o][i], SNt[i], t);
            mNt[o] = t; mN[o] = rec[PG_CN + 10 + o];
            rt[PG_CN + 10 + o] = t; eNt[10 + o] = SNt[o];
        }
        const float nv[2] = {rec[PG_CN + 23], rec[PG_CN + 24]};
        float m[PSI_D], hid[PSI_D];
        uint32_t hm;
        neumann_mlp(hi, mN, prb, nv, m, hm, hid);
        float hidt[PSI_D], mt[PSI_D];
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) {
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < PSI_D; ++i) t = fmaf(cW.un_W1[o][i], hdi[i], t);
#pragma unroll
            for (int i = 0; i < PSI_D; ++i) t = fmaf(cW.un_W1[o][PSI_D + i], mNt[i], t);
            hidt[o] = ((hm >> o) & 1u) ? t : 0.f;
            rt[PG_HIDN + o] = hidt[o];
        }
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) {
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < PSI_D; ++q) t = fmaf(cW.un_W2[o][q], hidt[q], t);
            mt[o] = t;                                   // r = m on Neumann rows
        }
        float rhat_t[PSI_D], rbar_t[PSI_D];
        ln_tan(rhat, rstd, mt, rbar, yi, rhat_t, rbar_t);
        float tbt[PSI_D], mpbt[PSI_D];
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) {
            rt[PG_RHAT + o] = rhat_t[o];
            rt[PG_MBN + o] = rbar_t[o];
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < PSI_D; ++q) t = fmaf(cW.un_W2[q][o], rbar_t[q], t);
            tbt[o] = ((hm >> o) & 1u) ? t : 0.f;
            rt[PG_TBN + o] = tbt[o];
        }
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) {
            float b = 0.f;
#pragma unroll
            for (int o = 0; o < PSI_D; ++o) b = fmaf(cW.un_W1[o][PSI_D + i], tbt[o], b);
            mpbt[i] = b;
        }
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) {
            float t = 0.f;
#pragma unroll
            for (int o = 0; o < PSI_D; ++o) t = fmaf(cW.neu.W2[o][i], mpbt[o], t);
            eNt[i] = mpbt[i];
            eNt[30 + i] = t;
            eNt[20 + i] = t * cntN[i];
        }
        return;
    }
    float STt[PSI_D], SFt[PSI_D], cntT[PSI_D], cntF[PSI_D];
    pgrad_list_tan<0>(G.T, h, hdot, node, hi, hdi, STt, cntT);
    pgrad_list_tan<1>(G.F, h, hdot, node, hi, hdi, SFt, cntF);
    float* eTt = rt + PG_EDGE;
    float* eFt = rt + PG_EDGE + 70;
    float mT[PSI_D], mF[PSI_D], mTt[PSI_D], mFt[PSI_D];
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) {
            a = fmaf(cW.to.W2[o][i], STt[i], a);
            b = fmaf(cW.from.W2[o][i], SFt[i], b);
        }
        mTt[o] = a; mFt[o] = b;
        mT[o] = rec[PG_C + 10 + o]; mF[o] = rec[PG_C + 20 + o];
        rt[PG_C + 10 + o] = a; rt[PG_C + 20 + o] = b;
        eTt[10 + o] = STt[o]; eFt[10 + o] = SFt[o];
    }
    const float alpha = gate<PRB>(hi, mT, mF, prb);
    float m[PSI_D], hid[PSI_D];
    uint32_t hm;
    update_mlp<PRB>(hi, mT, mF, prb, m, hm, hid);
    // gate and update MLP tangents
    float st = 0.f;
#pragma unroll
    for (int i = 0; i < PSI_D; ++i) {
        st = fmaf(cW.gate_w[i], hdi[i], st);
        st = fmaf(cW.gate_w[PSI_D + i], mTt[i], st);
        st = fmaf(cW.gate_w[2 * PSI_D + i], mFt[i], st);
    }
    const float dsig = alpha * (1.0f - alpha);
    const float alpha_t = dsig * st;
    float hidt[PSI_D], mt[PSI_D], rt_pre[PSI_D];
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) {
            t = fmaf(cW.up_W1[o][i], hdi[i], t);
            t = fmaf(cW.up_W1[o][PSI_D + i], mTt[i], t);
            t = fmaf(cW.up_W1[o][2 * PSI_D + i], mFt[i], t);
        }
        hidt[o] = ((hm >> o) & 1u) ? t : 0.f;
        rt[PG_HID + o] = hidt[o];
    }
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < PSI_D; ++q) t = fmaf(cW.up_W2[o][q], hidt[q], t);
        mt[o] = t;
        rt_pre[o] = hdi[o] + alpha_t * m[o] + alpha * t;           // r = h + α m
    }
    float rhat_t[PSI_D], rbar_t[PSI_D];
    ln_tan(rhat, rstd, rt_pre, rbar, yi, rhat_t, rbar_t);
    float abar = 0.f, abar_t = 0.f;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        abar = fmaf(rbar[o], m[o], abar);
        abar_t = fmaf(rbar_t[o], m[o], abar_t);
        abar_t = fmaf(rbar[o], mt[o], abar_t);
    }
    const float sbar_t = abar_t * dsig + abar * (1.0f - 2.0f * alpha) * alpha_t;
    rt[PG_SB] = sbar_t;
    float mbt[PSI_D], tbt[PSI_D];
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        rt[PG_RHAT + o] = rhat_t[o];
        mbt[o] = alpha_t * rbar[o] + alpha * rbar_t[o];
        rt[PG_MB + o] = mbt[o];
    }
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < PSI_D; ++q) t = fmaf(cW.up_W2[q][o], mbt[q], t);
        tbt[o] = ((hm >> o) & 1u) ? t : 0.f;
        rt[PG_TB + o] = tbt[o];
    }
    float mTbt[PSI_D], mFbt[PSI_D];
#pragma unroll
    for (int i = 0; i < PSI_D; ++i) {
        float b = sbar_t * cW.gate_w[PSI_D + i], c = sbar_t * cW.gate_w[2 * PSI_D + i];
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) {
            b = fmaf(cW.up_W1[o][PSI_D + i], tbt[o], b);
            c = fmaf(cW.up_W1[o][2 * PSI_D + i], tbt[o], c);
        }
        mTbt[i] = b; mFbt[i] = c;
    }
#pragma unroll
    for (int i = 0; i < PSI_D; ++i) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) {
            a = fmaf(cW.to.W2[o][i], mTbt[o], a);
            b = fmaf(cW.from.W2[o][i], mFbt[o], b);
        }
        eTt[i] = mTbt[i]; eTt[30 + i] = a; eTt[20 + i] = a * cntT[i];
        eFt[i] = mFbt[i]; eFt[30 + i] = b; eFt[20 + i] = b * cntF[i];
    }
}

template <int KIND>
__device__ __noinline__ void pgrad_node_outline(const GraphDev& G, const VjpCacheDev& C, const float* __restrict__ h, const float* __restrict__ y,
                                                const float* __restrict__ acc, int node, float* rec) {
    pgrad_node<KIND>(G, C, h, y, acc, node, rec);
}

// PASS 0: S̄' of every node into Sb_t (the planar padded [2][N][12] layout of VjpCacheDev::Sb).  PASS 1: accumulate.
template <int KIND>
__global__ void __launch_bounds__(PG_NODES)
k_pgrad_tan(GraphDev G, VjpCacheDev C, const float* __restrict__ h, const float* __restrict__ hdot, const float* __restrict__ y,
            const float* __restrict__ acc, const float* __restrict__ acc_t, const int* __restrict__ tab_y, const int* __restrict__ tab_x,
            int n_tab, float* __restrict__ partial, int num_batches, float* __restrict__ Sb_t, const int PASS) {
    extern __shared__ float rec[];                         // [PG_NODES][PG_PITCH] record, then [PG_NODES][PG_PITCH] tangent
    float* rect = rec + PG_NODES * PG_PITCH;
    float a[PG_MAX_PER_THREAD];
    int ty[PG_MAX_PER_THREAD], tx[PG_MAX_PER_THREAD];
    if (PASS == 1) {
#pragma unroll
        for (int k = 0; k < PG_MAX_PER_THREAD; ++k) {
            a[k] = 0.f;
            const int p = threadIdx.x + PG_NODES * k;
            ty[k] = (p < n_tab) ? tab_y[p] : 0;
            tx[k] = (p < n_tab) ? tab_x[p] : 0;
        }
    }
    for (int batch = blockIdx.x; batch < num_batches; batch += gridDim.x) {
        const int node = batch * PG_NODES + threadIdx.x;
        float* r = rec + threadIdx.x * PG_PITCH;
        float* rt = rect + threadIdx.x * PG_PITCH;
        pgrad_node_outline<KIND>(G, C, h, y, acc, node, r);
        pgrad_node_tan<KIND>(G, C, h, hdot, PASS == 1 ? acc_t : nullptr, node, r, rt);
        if (PASS == 0) {
            if (node < G.n_compute) {
                const bool neu = (KIND == KIND_MIXED) && (G.tag[node] & 2) && !(G.tag[node] & 1);
                float s0[PSI_D], s1[PSI_D];
#pragma unroll
                for (int o = 0; o < PSI_D; ++o) {
                    s0[o] = neu ? 0.f : rt[PG_EDGE + 30 + o];
                    s1[o] = neu ? rt[PG_EDGE + 140 + 30 + o] : rt[PG_EDGE + 70 + 30 + o];
                }
                store_row12(Sb_t, node, s0);
                store_row12(Sb_t + (int64_t)G.N * PSI_QPITCH, node, s1);
            }
            continue;
        }
        __syncthreads();
        for (int nd = 0; nd < PG_NODES; ++nd) {
            const float* p = rec + nd * PG_PITCH;
            const float* q = rect + nd * PG_PITCH;
#pragma unroll
            for (int k = 0; k < PG_MAX_PER_THREAD; ++k) a[k] = fmaf(q[ty[k]], p[tx[k]], fmaf(p[ty[k]], q[tx[k]], a[k]));
        }
        __syncthreads();
    }
    if (PASS == 1) {
#pragma unroll
        for (int k = 0; k < PG_MAX_PER_THREAD; ++k) {
            const int p = threadIdx.x + PG_NODES * k;
            if (p < n_tab) partial[(int64_t)blockIdx.x * n_tab + p] = a[k];
        }
    }
}
