// vjp.cuh — transpose-Jacobian product  y ↦ (∂f/∂h at H*)ᵀ y  at a frozen linearisation point.
//
// Replaces the full autograd graph walk the reference performs once per adjoint iteration
// (dirichlet/psignn/model.py:214: autograd.grad(new_H_star, H_star, y, retain_graph=True)).
//
// The scatters of reverse-mode autograd (index_select backward = float atomics on CUDA) are turned
// into gathers over the *other* adjacency list, so the result is deterministic and atomic-free:
//
//   prepare (once per backward solve): per node — LayerNorm statistics r̂, 1/σ, gate α, update output m,
//       hidden ReLU mask, and per destination the ReLU-activity counts of its incoming edge messages;
//       per list slot — the "cross" ReLU mask of the same edge seen from the neighbour's side
//       (slot of edge (r,c) in the row list of r holds the mask of Phi_to at destination c, and vice versa).
//   phase A (per node, no edge loop):  LayerNorm / gate / update-MLP backward → node-local part D_i and
//       S̄_i = W2ᵀ·m̄p_i for each direction; destination-side edge term collapses to W1iᵀ(S̄_i ⊙ count_i).
//   phase B (per node j, gather):  (Jᵀy)_j = D_j + W1j_toᵀ Σ_{(j,c)} S̄to_c ⊙ mask + W1j_fromᵀ Σ_{(r,j)} S̄from_r ⊙ mask.
//
// tests/restructured_math.py holds the torch prototype of exactly this algebra, checked against autograd.
#pragma once
#include "layer.cuh"

__device__ __forceinline__ uint32_t relu_bits(const float (&z)[PSI_D]) {
    uint32_t m = 0;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) m |= (z[o] > 0.f) ? (1u << o) : 0u;
    return m;
}

// own-direction aggregate with activity counts; optionally writes the cross mask of every slot.
// OWN   : edge MLP used at this destination (0 to, 1 from, 2 neumann) — only evaluated if `active`
// cross : for list T (slots = edges (j → i)), the neighbour j is the destination of the from/neumann MLP;
//         for list F, the neighbour is the destination of the to MLP.
template <int KIND, int LIST /*0 = T, 1 = F*/, int OWN>
__device__ __forceinline__ void prepare_list(const GraphDev& G, const SellDev& L, const float* __restrict__ h, int node,
                                             const float (&hi)[PSI_D], bool active, float (&mp)[PSI_D], float (&cnt)[PSI_D]) {
    const EdgeMLP& W = edge_mlp<OWN>();
    float P[PSI_D], S[PSI_D];
    edge_pre<OWN>(hi, P);
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) { S[o] = 0.f; cnt[o] = 0.f; }
    const int64_t base = L.slice_off[node >> 5];
    const int width = (int)((L.slice_off[(node >> 5) + 1] - base) >> 5);
    const int64_t col0 = base + (node & 31);
    int deg = 0;
    for (int t = 0; t < width; ++t) {
        const int64_t slot = col0 + (int64_t)t * 32;
        const int4 rec = __ldg(L.recs + slot);
        if (rec.x >= 0) {
            float hj[PSI_D], z[PSI_D], Qv[PSI_D];
            load_row(h, rec.x, hj);
            if (active) {
                edge_q<OWN>(hj, Qv);
                edge_z<OWN, 3>(P, Qv, rec, z);
#pragma unroll
                for (int o = 0; o < PSI_D; ++o) {
                    S[o] += fmaxf(z[o], 0.f);
                    cnt[o] += (z[o] > 0.f) ? 1.f : 0.f;
                }
                ++deg;
            }
            // cross mask: same edge, neighbour as destination, this node as source
            const uint8_t tj = G.tag[rec.x];
            uint32_t xm = 0;
            float Pj[PSI_D];
            if (LIST == 0) {                       // neighbour j = row: destination of Phi_from / phi_neumann
                if (!(tj & 1)) {
                    if (KIND == KIND_MIXED && (tj & 2)) {
                        edge_pre<2>(hj, Pj);
                        edge_q<2>(hi, Qv);
                        edge_z<2, 3>(Pj, Qv, rec, z);
                        xm = relu_bits(z) | (1u << 10);
                    } else {
                        edge_pre<1>(hj, Pj);
                        edge_q<1>(hi, Qv);
                        edge_z<1, 3>(Pj, Qv, rec, z);
                        xm = relu_bits(z);
                    }
                }
            } else {                               // neighbour j = col: destination of Phi_to
                if (!(tj & 1) && !(KIND == KIND_MIXED && (tj & 2))) {
                    edge_pre<0>(hj, Pj);
                    edge_q<0>(hi, Qv);
                    edge_z<0, 3>(Pj, Qv, rec, z);
                    xm = relu_bits(z);
                }
            }
            reinterpret_cast<int2*>(L.xmask)[slot] = make_int2(rec.x, (int)xm);
        }
    }
    const float fdeg = (float)deg;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        float t_ = fdeg * W.b2[o];
#pragma unroll
        for (int i = 0; i < PSI_D; ++i) t_ = fmaf(W.W2[o][i], S[i], t_);
        mp[o] = t_;
    }
}

template <int KIND>
__global__ void __launch_bounds__(PSI_NODE_BLOCK)
k_vjp_prepare(GraphDev G, VjpCacheDev C, const float* __restrict__ h) {
    const int node = blockIdx.x * PSI_NODE_BLOCK + threadIdx.x;
    if (node >= G.n_compute) return;                  // mesh partition: the cache of a ghost row lives with its owner
    constexpr int PRB = (KIND == KIND_MIXED) ? 3 : 2;
    float hi[PSI_D];
    load_row(h, node, hi);
    const uint8_t tg = G.tag[node];
    const bool dir = tg & 1;
    const bool neu = (KIND == KIND_MIXED) && (tg & 2);
    const bool interior = !dir && !neu;
    float mT[PSI_D], mF[PSI_D], cT[PSI_D], cF[PSI_D], cN[PSI_D];
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) cN[o] = 0.f;
    prepare_list<KIND, 0, 0>(G, G.T, h, node, hi, interior, mT, cT);
    if (neu) {
        float dummy[PSI_D];
        prepare_list<KIND, 1, 2>(G, G.F, h, node, hi, true, mF, cN);     // mF holds mp_neumann
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) { cF[o] = 0.f; dummy[o] = 0.f; }
        (void)dummy;
    } else {
        prepare_list<KIND, 1, 1>(G, G.F, h, node, hi, interior, mF, cF);
    }
    float r[PSI_D], m[PSI_D], alpha = 0.f;
    uint32_t hm = 0;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) { r[o] = 0.f; m[o] = 0.f; }
    if (!dir) {
        float prb[3];
        load_prb<PRB>(G, node, prb);
        if (neu) {
            const float nv[2] = {__ldg(G.nrm + 2 * (int64_t)node), __ldg(G.nrm + 2 * (int64_t)node + 1)};
            neumann_mlp(hi, mF, prb, nv, m, hm);
            hm <<= 10;
#pragma unroll
            for (int o = 0; o < PSI_D; ++o) r[o] = m[o];
        } else {
            alpha = gate<PRB>(hi, mT, mF, prb);
            update_mlp<PRB>(hi, mT, mF, prb, m, hm);
#pragma unroll
            for (int o = 0; o < PSI_D; ++o) r[o] = fmaf(alpha, m[o], hi[o]);
        }
    }
    float outv[PSI_D], rhat[PSI_D], rstd;
    layer_norm10(r, outv, rhat, rstd);
    store_row(C.rhat, node, rhat);
    store_row(C.m, node, m);
    C.rstd[node] = rstd;
    C.alpha[node] = alpha;
    C.nmask[node] = hm;
    store_row(C.cnt, (int64_t)node * 3 + 0, cT);
    store_row(C.cnt, (int64_t)node * 3 + 1, cF);
    store_row(C.cnt, (int64_t)node * 3 + 2, cN);
}

// out[i] += Σ_o W[o][i]·v[o]   (transpose application of a d×d block)
#define PSI_TMATVEC(Wmat, v, out)                                             \
    _Pragma("unroll") for (int i_ = 0; i_ < PSI_D; ++i_) {                    \
        float t_ = out[i_];                                                   \
        _Pragma("unroll") for (int o_ = 0; o_ < PSI_D; ++o_) t_ = fmaf(Wmat[o_][i_], v[o_], t_); \
        out[i_] = t_;                                                         \
    }

template <int KIND>
__global__ void __launch_bounds__(PSI_NODE_BLOCK)
k_vjp_phase_a(GraphDev G, VjpCacheDev C, const float* __restrict__ y, const int* __restrict__ done) {
    if (done != nullptr && *done) return;
    const int node = blockIdx.x * PSI_NODE_BLOCK + threadIdx.x;
    if (node >= G.n_compute) return;                  // mesh partition: S̄ of the ghost rows arrives from their owners
    constexpr int PRB = (KIND == KIND_MIXED) ? 3 : 2;
    const uint8_t tg = G.tag[node];
    float D[PSI_D], SbT[PSI_D], SbF[PSI_D];
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) { D[o] = 0.f; SbT[o] = 0.f; SbF[o] = 0.f; }
    if (!(tg & 1)) {
        float yi[PSI_D], rhat[PSI_D], rbar[PSI_D];
        load_row(y, node, yi);
        load_row(C.rhat, node, rhat);
        const float rstd = C.rstd[node];
        // LayerNorm backward: r̄ = rstd·(ĝ − mean(ĝ) − r̂·mean(ĝ⊙r̂)),  ĝ = γ⊙ȳ
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
        const uint32_t nm = C.nmask[node];
        if (KIND == KIND_MIXED && (tg & 2)) {
            // r = update_neumann(cat[h, mp_neu, prb, n̂]) — no residual connection
            float hb[PSI_D], mpb[PSI_D];
#pragma unroll
            for (int o = 0; o < PSI_D; ++o) {
                float t = 0.f;
#pragma unroll
                for (int q = 0; q < PSI_D; ++q) t = fmaf(cW.un_W2[q][o], rbar[q], t);
                hb[o] = ((nm >> (10 + o)) & 1u) ? t : 0.f;
            }
#pragma unroll
            for (int i = 0; i < PSI_D; ++i) {
                float a = 0.f, b = 0.f;
#pragma unroll
                for (int o = 0; o < PSI_D; ++o) {
                    a = fmaf(cW.un_W1[o][i], hb[o], a);
                    b = fmaf(cW.un_W1[o][PSI_D + i], hb[o], b);
                }
                D[i] = a;
                mpb[i] = b;
            }
            // S̄N = W2_neuᵀ·m̄p ; destination-side term W1i_neuᵀ(S̄N ⊙ count)
            float cn[PSI_D], sc[PSI_D];
            load_row(C.cnt, (int64_t)node * 3 + 2, cn);
#pragma unroll
            for (int i = 0; i < PSI_D; ++i) {
                float t = 0.f;
#pragma unroll
                for (int o = 0; o < PSI_D; ++o) t = fmaf(cW.neu.W2[o][i], mpb[o], t);
                SbF[i] = t;
                sc[i] = t * cn[i];
            }
            PSI_TMATVEC(cW.neu.W1i, sc, D);
        } else {
            float m[PSI_D];
            load_row(C.m, node, m);
            const float alpha = C.alpha[node];
            float abar = 0.f;
#pragma unroll
            for (int o = 0; o < PSI_D; ++o) abar = fmaf(rbar[o], m[o], abar);
            const float sbar = abar * alpha * (1.0f - alpha);
            float hb[PSI_D];
#pragma unroll
            for (int o = 0; o < PSI_D; ++o) {
                float t = 0.f;
#pragma unroll
                for (int q = 0; q < PSI_D; ++q) t = fmaf(cW.up_W2[q][o], alpha * rbar[q], t);
                hb[o] = ((nm >> o) & 1u) ? t : 0.f;
            }
            float mTb[PSI_D], mFb[PSI_D];
#pragma unroll
            for (int i = 0; i < PSI_D; ++i) {
                float a = sbar * cW.gate_w[i], b = sbar * cW.gate_w[PSI_D + i], c = sbar * cW.gate_w[2 * PSI_D + i];
#pragma unroll
                for (int o = 0; o < PSI_D; ++o) {
                    a = fmaf(cW.up_W1[o][i], hb[o], a);
                    b = fmaf(cW.up_W1[o][PSI_D + i], hb[o], b);
                    c = fmaf(cW.up_W1[o][2 * PSI_D + i], hb[o], c);
                }
                D[i] = rbar[i] + a;
                mTb[i] = b;
                mFb[i] = c;
            }
            float ct[PSI_D], cf[PSI_D], sct[PSI_D], scf[PSI_D];
            load_row(C.cnt, (int64_t)node * 3 + 0, ct);
            load_row(C.cnt, (int64_t)node * 3 + 1, cf);
#pragma unroll
            for (int i = 0; i < PSI_D; ++i) {
                float a = 0.f, b = 0.f;
#pragma unroll
                for (int o = 0; o < PSI_D; ++o) {
                    a = fmaf(cW.to.W2[o][i], mTb[o], a);
                    b = fmaf(cW.from.W2[o][i], mFb[o], b);
                }
                SbT[i] = a; sct[i] = a * ct[i];
                SbF[i] = b; scf[i] = b * cf[i];
            }
            PSI_TMATVEC(cW.to.W1i, sct, D);
            PSI_TMATVEC(cW.from.W1i, scf, D);
        }
        (void)PRB;
    }
    store_row(C.Dloc, node, D);
    store_row12(C.Sb, node, SbT);                     // planar [2][N][12]: the gathers of phase B read contiguous 16-byte aligned rows
    store_row12(C.Sb, (int64_t)G.N + node, SbF);
}

// phase B's per-record work is ten selects and adds: the staged cooperative gather costs more instructions than it saves cache
// wavefronts there (measured: 22.0 µs vs 17.8 µs at C3, 108 µs vs 102 µs at C5), so its rows are gathered directly
#ifndef PSI_VJP_GATHER_DIRECT
#define PSI_VJP_GATHER_DIRECT 1
#endif
#if PSI_VJP_GATHER_DIRECT
#define PSI_WALK_B(recs, width, src, lane, W, M, key, body) walk_direct(recs, width, src, lane, key, body)
#else
#define PSI_WALK_B(recs, width, src, lane, W, M, key, body) walk_ring(recs, width, src, lane, W, M, key, body)
#endif

template <int KIND, bool EPI>
__global__ void __launch_bounds__(PSI_NODE_BLOCK, PSI_OP_MIN_CTAS)
k_vjp_phase_b(GraphDev G, VjpCacheDev C, const float* __restrict__ y, const float* __restrict__ grad,
              float* __restrict__ out, SolverEpi E, float* __restrict__ acc_out /* optional [N][30]: the gathered sums themselves (pgrad.cuh) */) {
    __shared__ WarpStage stage[PSI_NODE_BLOCK / 32];
    __shared__ float smem[2 * PSI_NODE_BLOCK / 32];
    if (EPI && *E.done) return;
    const int node = blockIdx.x * PSI_NODE_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = node < G.n_compute;
    float yi[PSI_D], res[PSI_D];
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) { yi[o] = 0.f; res[o] = 0.f; }
    if ((node & ~31) < G.n_compute) {                  // warp-uniform
        float accT[PSI_D], accF[PSI_D], accN[PSI_D];
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) { accT[o] = 0.f; accF[o] = 0.f; accN[o] = 0.f; }
        CoopMap M;
        coop_map(lane, M);
        WarpStage& W = stage[threadIdx.x >> 5];
        const int slice = node >> 5;
        const int N = G.N;
        // Both walks gather S̄ of the destination over the *other* list ({neighbour, cross mask} pairs written by k_vjp_prepare; a
        // zero mask — padding slots, Dirichlet destinations — selects nothing): the scatter of reverse-mode autograd as a gather.
        // row list of this node: edges (node, c) — node is the source of Phi_to messages into c
        {
            const int64_t base = G.F.slice_off[slice];
            const int width = (int)((G.F.slice_off[slice + 1] - base) >> 5);
            PSI_WALK_B(reinterpret_cast<const int2*>(G.F.xmask) + base, width, C.Sb, lane, W, M,
                      [&](const int2& jm) { return jm.y ? jm.x : 0; },
                      [&](const int2& jm, const f2 (&q2)[PSI_D / 2]) {
                          float sb[PSI_D];
                          unpack10(q2, sb);
                          const uint32_t xm = (uint32_t)jm.y;
#pragma unroll
                          for (int o = 0; o < PSI_D; ++o) accT[o] += ((xm >> o) & 1u) ? sb[o] : 0.f;
                      });
        }
        // column list: edges (r, node) — node is the source of Phi_from / phi_neumann messages into r
        {
            const int64_t base = G.T.slice_off[slice];
            const int width = (int)((G.T.slice_off[slice + 1] - base) >> 5);
            PSI_WALK_B(reinterpret_cast<const int2*>(G.T.xmask) + base, width, C.Sb, lane, W, M,
                      [&](const int2& jm) { return jm.y ? N + jm.x : 0; },
                      [&](const int2& jm, const f2 (&q2)[PSI_D / 2]) {
                          float sb[PSI_D];
                          unpack10(q2, sb);
                          const uint32_t xm = (uint32_t)jm.y;
                          if (KIND == KIND_MIXED) {
                              const bool neu = xm & (1u << 10);
#pragma unroll
                              for (int o = 0; o < PSI_D; ++o) {
                                  const float v = ((xm >> o) & 1u) ? sb[o] : 0.f;
                                  accN[o] += neu ? v : 0.f;
                                  accF[o] += neu ? 0.f : v;
                              }
                          } else {
#pragma unroll
                              for (int o = 0; o < PSI_D; ++o) accF[o] += ((xm >> o) & 1u) ? sb[o] : 0.f;
                          }
                      });
        }
        if (valid && acc_out != nullptr) {
            store_row(acc_out, (int64_t)node * 3 + 0, accT);
            store_row(acc_out, (int64_t)node * 3 + 1, accF);
            store_row(acc_out, (int64_t)node * 3 + 2, accN);
        }
        if (valid) {
            load_row_rw(C.Dloc, node, res);
            PSI_TMATVEC(cW.to.W1j, accT, res);
            PSI_TMATVEC(cW.from.W1j, accF, res);
            if (KIND == KIND_MIXED) { PSI_TMATVEC(cW.neu.W1j, accN, res); }
            if (grad != nullptr) {
                float gi[PSI_D];
                load_row(grad, node, gi);
#pragma unroll
                for (int o = 0; o < PSI_D; ++o) res[o] += gi[o];
            }
            if (EPI) load_row(y, node, yi);
            else store_row(out, node, res);
        }
    }
    if (EPI) solver_epilogue(E, node, valid, yi, res, smem);
}
