// baseline_bwd.cuh — backward of ONE unrolled layer of the DSS / DSGPS baselines (layer kinds 2, 3, 4):
//     h̄ = (∂f/∂h)ᵀ ȳ   and   θ̄ = (∂f/∂θ)ᵀ ȳ    at the layer's own input h.
//
// Replaces the autograd walk through one unrolled step of the reference's training forward
// (dirichlet/dss/model.py:83-91, dirichlet/dsgps/model.py:143-163, mixed/dsgps/model.py:76-97; its index_select / scatter backward
// are float atomics on CUDA).  Same restructuring as vjp.cuh + pgrad.cuh — scatters become gathers over the other adjacency list,
// every parameter gradient is Σ_nodes rec[ty]·rec[tx] in a fixed order — but the linearisation point is used once (each unrolled
// layer has its own input), so nothing is cached: the forward intermediates and the ReLU masks are recomputed where they are needed,
// with the canonical rounding chains of layer.cuh (edge_pre / edge_q / edge_z), so every mask equals the forward's bit for bit.
//
//   pass 0 (k_bl_pass, per node i):   recompute the aggregation and the node update, walk it backwards → node-local part D_i of h̄
//                                     (incl. the destination side W1iᵀ(S̄ ⊙ count) of the edge MLPs) and S̄_i = W2ᵀ·m̄p_i per direction
//   gather (k_bl_gather, per node j): acc_X,j = Σ_{edges leaving j} mask_e ⊙ S̄_X,dst(e);  h̄_j = D_j + Σ_X W1j_Xᵀ acc_X,j
//   pass 1 (k_bl_pass):               the node record again (now with acc) and the table sums → θ̄
//
// The node record reuses the layout of pgrad.cuh (psi_pgrad_layout) with these meanings for the baseline kinds:
//   PG_C    c  = [h, ΣΦ→, ΣΦ←, second member]            PG_CN  cN = [h, ΣΦ_neumann, second member, normal]
//   PG_TB   DSS: t̄ = hidden cotangent ⊙ mask              DSGPS: ā  (cotangent of the z-gate pre-activation)
//   PG_MB   DSS: m̄ = α·ȳ                                   DSGPS: b̄  (r-gate pre-activation)
//   PG_YB                                                   DSGPS: ē  (correction pre-tanh)
//   PG_HID  DSS: hidden of Ψ                                PG_RHAT DSGPS: r ⊙ h (first 10 inputs of the correction)
//   PG_EDGE, PG_ACC, PG_MBN, PG_HIDN, PG_TBN as in pgrad.cuh.
#pragma once
#include "pgrad.cuh"

#define BL_MAX_PER_THREAD 40             // table entries ≤ 64·40 = 2560 (mixed DSGPS layer: 2 440)

// own-direction aggregate of one destination: S = Σ relu(z), activity counts, mask-weighted attribute sums, degree, message
template <int OWN, int ATTR>
__device__ __forceinline__ void bl_list(const SellDev& L, const float* __restrict__ h, int node, const float (&hi)[PSI_D], float* rec_edge,
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
        edge_z<OWN, ATTR>(P, Qv, r, z);
        const float a0 = __int_as_float(r.y), a1 = (ATTR > 1) ? __int_as_float(r.z) : 0.f, a2 = (ATTR > 2) ? __int_as_float(r.w) : 0.f;
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
        rec_edge[20 + o] = cnt[o];            // becomes zs = S̄ ⊙ cnt in bl_edge_tail
        rec_edge[40 + 3 * o] = A[o][0]; rec_edge[41 + 3 * o] = A[o][1]; rec_edge[42 + 3 * o] = A[o][2];
    }
}

// message cotangent m̄p of one direction → S̄ = W2ᵀ m̄p, zs = S̄ ⊙ count, D += W1iᵀ zs
template <int OWN>
__device__ __forceinline__ void bl_edge_tail(const float (&mpb)[PSI_D], float* e, float (&D)[PSI_D]) {
    const EdgeMLP& W = edge_mlp<OWN>();
    float zs[PSI_D];
#pragma unroll
    for (int i = 0; i < PSI_D; ++i) {
        float t = 0.f;
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) t = fmaf(W.W2[o][i], mpb[o], t);
        e[i] = mpb[i];
        e[30 + i] = t;
        zs[i] = t * e[20 + i];
        e[20 + i] = zs[i];
    }
    PSI_TMATVEC(W.W1i, zs, D);
}

// forward intermediates of the GRU-style DSGPS update, same operation order as dsgps_update (layer.cuh)
template <int PRB>
__device__ __forceinline__ void dsgps_parts(const float (&hi)[PSI_D], const float (&mT)[PSI_D], const float (&mF)[PSI_D], const float (&prb)[3],
                                            float (&zk)[PSI_D], float (&rk)[PSI_D], float (&th)[PSI_D]) {
    float c[33];
#pragma unroll
    for (int i = 0; i < PSI_D; ++i) { c[i] = hi[i]; c[PSI_D + i] = mT[i]; c[2 * PSI_D + i] = mF[i]; }
#pragma unroll
    for (int i = 0; i < 3; ++i) c[30 + i] = prb[i];
    {
        f2 az[PSI_D / 2], ar[PSI_D / 2];
        bias2(cW.gz_b, az);
        bias2(cW.gr_b, ar);
        mv2<30 + PRB>(cWT.gzT, c, az);
        mv2<30 + PRB>(cWT.grT, c, ar);
        float a[PSI_D], b[PSI_D];
        unpack10(az, a);
        unpack10(ar, b);
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) { zk[o] = sigmoidf_acc(a[o]); rk[o] = sigmoidf_acc(b[o]); }
    }
#pragma unroll
    for (int i = 0; i < PSI_D; ++i) c[i] = rk[i] * hi[i];
    f2 ac[PSI_D / 2];
    bias2(cW.gc_b, ac);
    mv2<30 + PRB>(cWT.gcT, c, ac);
    float a[PSI_D];
    unpack10(ac, a);
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) th[o] = tanhf(a[o]);
}

// fills rec[0 .. PG_REC) of one node and returns the node-local part D of h̄.  Inlined on purpose: as a __noinline__ function (GraphDev
// passed by reference through a local copy) the mixed-kind instance faulted in its Neumann list walk on B200 (nvcc 12.9) while the
// inlined one is correct — bisected with build variants, cause not established.
template <int KIND>
__device__ __forceinline__ void bl_node(const GraphDev& G, const float* __restrict__ h, const float* __restrict__ y, const float* __restrict__ acc,
                                     int node, float* rec, float (&D)[PSI_D]) {
    constexpr int PRB = KindTraits<KIND>::PRB;
    constexpr int ATTR = KindTraits<KIND>::ATTR;
    for (int i = 0; i < PG_REC; ++i) rec[i] = 0.f;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) D[o] = 0.f;
    if (node >= G.n_compute) return;
    rec[PG_ONE] = 1.f;
    float hi[PSI_D];
    load_row(h, node, hi);
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) { rec[PG_C + o] = hi[o]; rec[PG_CN + o] = hi[o]; }
    if (acc != nullptr)
        for (int i = 0; i < 30; ++i) rec[PG_ACC + i] = acc[(int64_t)node * 30 + i];
    const int cls = node_class<KIND>(G, node, G.n_compute);
    if (cls == 1) return;                                 // Dirichlet row: f = h0 (no parameter, no dependence on h); it still sends messages
    float yi[PSI_D], prb[3];
    load_row(y, node, yi);
    load_prb<PRB>(G, node, prb);
    if (KindTraits<KIND>::has_neumann && cls == 2) {
        // out = update_neumann(cat[h, ΣΦ_neumann, prb, n̂]) — no residual connection, no LayerNorm (mixed/dsgps/model.py:88-92)
        float mN[PSI_D], degN;
        float* eN = rec + PG_EDGE + 140;
        bl_list<2, ATTR>(G.F, h, node, hi, eN, degN, mN);
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
            rec[PG_MBN + o] = yi[o]; rec[PG_HIDN + o] = hid[o];
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < PSI_D; ++q) t = fmaf(cW.un_W2[q][o], yi[q], t);
            tb[o] = ((hm >> o) & 1u) ? t : 0.f;
            rec[PG_TBN + o] = tb[o];
        }
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int o = 0; o < PSI_D; ++o) {
                a = fmaf(cW.un_W1[o][i], tb[o], a);
                b = fmaf(cW.un_W1[o][PSI_D + i], tb[o], b);
            }
            D[i] = a;
            mpb[i] = b;
        }
        bl_edge_tail<2>(mpb, eN, D);
        return;
    }
    float mT[PSI_D], mF[PSI_D], degT, degF;
    float* eT = rec + PG_EDGE;
    float* eF = rec + PG_EDGE + 70;
    bl_list<0, ATTR>(G.T, h, node, hi, eT, degT, mT);
    bl_list<1, ATTR>(G.F, h, node, hi, eF, degF, mF);
    rec[PG_DEG + 0] = degT; rec[PG_DEG + 1] = degF;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) { rec[PG_C + 10 + o] = mT[o]; rec[PG_C + 20 + o] = mF[o]; }
    rec[PG_C + 30] = prb[0]; rec[PG_C + 31] = prb[1]; rec[PG_C + 32] = (PRB > 2) ? prb[2] : 0.f;
    float mTb[PSI_D], mFb[PSI_D];
    if (KIND == KIND_DSS) {
        // out = h + α·Ψ(c),  Ψ = W2·relu(W1·c + b1) + b2   (dirichlet/dss/model.py:113-119)
        float m[PSI_D], hid[PSI_D], tb[PSI_D];
        uint32_t hm;
        update_mlp<3>(hi, mT, mF, prb, m, hm, hid);
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) {
            rec[PG_MB + o] = cW.dss_alpha * yi[o];
            rec[PG_HID + o] = hid[o];
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < PSI_D; ++q) t = fmaf(cW.up_W2[q][o], cW.dss_alpha * yi[q], t);
            tb[o] = ((hm >> o) & 1u) ? t : 0.f;
            rec[PG_TB + o] = tb[o];
        }
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) {
            float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
            for (int o = 0; o < PSI_D; ++o) {
                a = fmaf(cW.up_W1[o][i], tb[o], a);
                b = fmaf(cW.up_W1[o][PSI_D + i], tb[o], b);
                c = fmaf(cW.up_W1[o][2 * PSI_D + i], tb[o], c);
            }
            D[i] = yi[i] + a;
            mTb[i] = b;
            mFb[i] = c;
        }
    } else {
        // out = h + z ⊙ tanh(C·[r ⊙ h, mT, mF, prb] + bc),  z = σ(Z·c + bz),  r = σ(R·c + br)   (dirichlet/dsgps/model.py:148-155)
        float zk[PSI_D], rk[PSI_D], th[PSI_D], ab[PSI_D], bb[PSI_D], eb[PSI_D];
        dsgps_parts<PRB>(hi, mT, mF, prb, zk, rk, th);
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) {
            eb[o] = yi[o] * zk[o] * (1.0f - th[o] * th[o]);
            ab[o] = yi[o] * th[o] * zk[o] * (1.0f - zk[o]);
            rec[PG_YB + o] = eb[o];
            rec[PG_TB + o] = ab[o];
            rec[PG_RHAT + o] = rk[o] * hi[o];
        }
        float c2h[PSI_D];
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) {
            float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
            for (int o = 0; o < PSI_D; ++o) {
                a = fmaf(cW.gc_W[o][i], eb[o], a);
                b = fmaf(cW.gc_W[o][PSI_D + i], eb[o], b);
                c = fmaf(cW.gc_W[o][2 * PSI_D + i], eb[o], c);
            }
            c2h[i] = a;
            mTb[i] = b;
            mFb[i] = c;
            bb[i] = a * hi[i] * rk[i] * (1.0f - rk[i]);
            rec[PG_MB + i] = bb[i];
        }
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) {
            float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
            for (int o = 0; o < PSI_D; ++o) {
                a = fmaf(cW.gz_W[o][i], ab[o], a);
                b = fmaf(cW.gz_W[o][PSI_D + i], ab[o], b);
                c = fmaf(cW.gz_W[o][2 * PSI_D + i], ab[o], c);
                a = fmaf(cW.gr_W[o][i], bb[o], a);
                b = fmaf(cW.gr_W[o][PSI_D + i], bb[o], b);
                c = fmaf(cW.gr_W[o][2 * PSI_D + i], bb[o], c);
            }
            D[i] = yi[i] + c2h[i] * rk[i] + a;
            mTb[i] += b;
            mFb[i] += c;
        }
    }
    bl_edge_tail<0>(mTb, eT, D);
    bl_edge_tail<1>(mFb, eF, D);
}

// PASS 0: D and S̄ of every node (planar padded [2][N][12] like VjpCacheDev::Sb).  PASS 1: table sums, one row of partials per CTA
// (fixed assignment of node batches to CTAs: deterministic).
template <int KIND>
__global__ void __launch_bounds__(PG_NODES)
k_bl_pass(GraphDev G, const float* __restrict__ h, const float* __restrict__ y, const float* __restrict__ acc, float* __restrict__ Dloc,
          float* __restrict__ Sb, const int* __restrict__ tab_y, const int* __restrict__ tab_x, int n_tab, float* __restrict__ partial,
          int num_batches, const int PASS) {
    extern __shared__ float rec[];                         // [PG_NODES][PG_PITCH]
    float a[BL_MAX_PER_THREAD];
    int ty[BL_MAX_PER_THREAD], tx[BL_MAX_PER_THREAD];
    if (PASS == 1) {
#pragma unroll
        for (int k = 0; k < BL_MAX_PER_THREAD; ++k) {
            a[k] = 0.f;
            const int p = threadIdx.x + PG_NODES * k;
            ty[k] = (p < n_tab) ? tab_y[p] : 0;
            tx[k] = (p < n_tab) ? tab_x[p] : 0;
        }
    }
    for (int batch = blockIdx.x; batch < num_batches; batch += gridDim.x) {
        const int node = batch * PG_NODES + threadIdx.x;
        float* r = rec + threadIdx.x * PG_PITCH;
        float D[PSI_D];
        bl_node<KIND>(G, h, y, PASS == 1 ? acc : nullptr, node, r, D);
        if (PASS == 0) {
            if (node < G.n_compute) {
                const int cls = node_class<KIND>(G, node, G.n_compute);
                float s0[PSI_D], s1[PSI_D];
#pragma unroll
                for (int o = 0; o < PSI_D; ++o) {
                    s0[o] = (cls == 0) ? r[PG_EDGE + 30 + o] : 0.f;
                    s1[o] = (cls == 0) ? r[PG_EDGE + 70 + 30 + o] : (cls == 2 ? r[PG_EDGE + 140 + 30 + o] : 0.f);
                }
                store_row(Dloc, node, D);
                store_row12(Sb, node, s0);
                store_row12(Sb + (int64_t)G.N * PSI_QPITCH, node, s1);
            }
            continue;
        }
        __syncthreads();
        for (int nd = 0; nd < PG_NODES; ++nd) {
            const float* q = rec + nd * PG_PITCH;
#pragma unroll
            for (int k = 0; k < BL_MAX_PER_THREAD; ++k) a[k] = fmaf(q[ty[k]], q[tx[k]], a[k]);
        }
        __syncthreads();
    }
    if (PASS == 1) {
#pragma unroll
        for (int k = 0; k < BL_MAX_PER_THREAD; ++k) {
            const int p = threadIdx.x + PG_NODES * k;
            if (p < n_tab) partial[(int64_t)blockIdx.x * n_tab + p] = a[k];
        }
    }
}

__device__ __forceinline__ void load_row12(const float* __restrict__ base, int64_t row, float (&v)[PSI_D]) {
    const float4* p = reinterpret_cast<const float4*>(base + row * PSI_QPITCH);
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; v[8] = c.x; v[9] = c.y;
}

// acc += mask(z of the edge at its destination `dst`, this node as the source) ⊙ S̄[dst]
template <int OWN, int ATTR>
__device__ __forceinline__ void bl_cross(const float* __restrict__ h, const float* __restrict__ sb_plane, int dst, const float (&Qsrc)[PSI_D],
                                         const int4& rec, float (&acc)[PSI_D]) {
    float hd[PSI_D], P[PSI_D], z[PSI_D], sb[PSI_D];
    load_row(h, dst, hd);
    edge_pre<OWN>(hd, P);
    edge_z<OWN, ATTR>(P, Qsrc, rec, z);
    load_row12(sb_plane, dst, sb);
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) acc[o] += (z[o] > 0.f) ? sb[o] : 0.f;
}

template <int KIND>
__global__ void __launch_bounds__(PSI_NODE_BLOCK)
k_bl_gather(GraphDev G, const float* __restrict__ h, const float* __restrict__ Dloc, const float* __restrict__ Sb, float* __restrict__ out,
            float* __restrict__ acc_out) {
    constexpr int ATTR = KindTraits<KIND>::ATTR;
    constexpr bool NEU = KindTraits<KIND>::has_neumann;
    const int node = blockIdx.x * PSI_NODE_BLOCK + threadIdx.x;
    if (node >= G.n_compute) return;
    float hj[PSI_D], Q0[PSI_D], Q1[PSI_D], Q2[PSI_D];
    load_row(h, node, hj);
    edge_q<0>(hj, Q0);
    edge_q<1>(hj, Q1);
    if (NEU) edge_q<2>(hj, Q2);
    float accT[PSI_D], accF[PSI_D], accN[PSI_D];
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) { accT[o] = 0.f; accF[o] = 0.f; accN[o] = 0.f; }
    const float* sb0 = Sb;
    const float* sb1 = Sb + (int64_t)G.N * PSI_QPITCH;
    const int slice = node >> 5;
    // row list of this node: edges (node, c) — node is the source of the Phi_to message into c
    {
        const int64_t base = G.F.slice_off[slice];
        const int width = (int)((G.F.slice_off[slice + 1] - base) >> 5);
        const int4* p = G.F.recs + base + (node & 31);
        for (int t = 0; t < width; ++t) {
            const int4 rec = __ldg(p + (int64_t)t * 32);
            if (rec.x < 0) continue;
            if (node_class<KIND>(G, rec.x, G.N) == 0) bl_cross<0, ATTR>(h, sb0, rec.x, Q0, rec, accT);
        }
    }
    // column list: edges (r, node) — node is the source of the Phi_from / phi_neumann message into r
    {
        const int64_t base = G.T.slice_off[slice];
        const int width = (int)((G.T.slice_off[slice + 1] - base) >> 5);
        const int4* p = G.T.recs + base + (node & 31);
        for (int t = 0; t < width; ++t) {
            const int4 rec = __ldg(p + (int64_t)t * 32);
            if (rec.x < 0) continue;
            const int cls = node_class<KIND>(G, rec.x, G.N);
            if (cls == 0) bl_cross<1, ATTR>(h, sb1, rec.x, Q1, rec, accF);
            else if (NEU && cls == 2) bl_cross<2, ATTR>(h, sb1, rec.x, Q2, rec, accN);
        }
    }
    float res[PSI_D];
    load_row(Dloc, node, res);
    PSI_TMATVEC(cW.to.W1j, accT, res);
    PSI_TMATVEC(cW.from.W1j, accF, res);
    if (NEU) { PSI_TMATVEC(cW.neu.W1j, accN, res); }
    store_row(out, node, res);
    store_row(acc_out, (int64_t)node * 3 + 0, accT);
    store_row(acc_out, (int64_t)node * 3 + 1, accF);
    store_row(acc_out, (int64_t)node * 3 + 2, accN);
}
