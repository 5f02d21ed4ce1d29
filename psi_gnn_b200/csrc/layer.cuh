// layer.cuh — the fused message-passing layer f_theta: a per-node pre-pass + one fused kernel per application.
//
//   k_layer_pre      Q_w[j] = W1j_w · h_j for the 2 (3) edge MLPs, once per node — the first edge layer is linear in the
//                    concatenation [h_i, h_j, a], so its h_j block is a per-SOURCE quantity: the per-edge work
//                    drops from 130 to 30 FMAs + 10 adds.
//   k_layer_forward  one thread owns one destination node (one warp = one 32-node slice of the SELL lists):
//     1. coalesced 16-byte edge records {j, a0, a1, a2}; every lane gathers the 48-byte padded row Q[j] of its neighbour with three
//        16-byte loads (two trips in flight).  The round-1 kernel gathered h[j] with five 8-byte loads per edge and was bound by L1
//        wavefronts (ncu: profiles/r02_a_operator.md); a cooperative shared-memory staged gather (walk_ring below) halves the
//        wavefronts again but costs more instructions than it saves once the arithmetic is packed.
//     The contractions are packed fp32x2 FMAs (FFMA2: two output channels per issue slot, bit-identical to scalar fma chains).
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

// ---- packed matrix-vector products ------------------------------------------------------------------------------------
// acc[q] (output channels 2q, 2q+1) += Σ_{i<NI} WT[i][2q..2q+1] · x[i]: the sequential fma chain over i of the scalar code, two
// output channels per FFMA2 (the weight pair is one 8-byte constant operand, x[i] a broadcast scalar)
template <int NI>
__device__ __forceinline__ void mv2(const float (*WT)[PSI_D], const float* x, f2 (&acc)[PSI_D / 2]) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const f2 xx = pk(x[i], x[i]);
#pragma unroll
        for (int q = 0; q < PSI_D / 2; ++q) acc[q] = ffma2(pk(WT[i][2 * q], WT[i][2 * q + 1]), xx, acc[q]);
    }
}
__device__ __forceinline__ void bias2(const float* b, f2 (&acc)[PSI_D / 2]) {
#pragma unroll
    for (int q = 0; q < PSI_D / 2; ++q) acc[q] = pk(b[2 * q], b[2 * q + 1]);
}

// ---- canonical rounding of the first edge layer -------------------------------------------------------------------
// Every kernel that needs z (forward, VJP-prepare own and cross masks, parameter gradients) uses exactly these three chains, so
// that a ReLU decision is the same wherever it is taken:
//   P[o] = b1[o] + Σ_i W1i[o][i]·h_dst[i]          (fma chain from the bias)
//   Q[o] =          Σ_i W1j[o][i]·h_src[i]          (fma chain from zero)
//   z[o] = fma(W1a[o][2], a2, fma(W1a[o][1], a1, fma(W1a[o][0], a0, P[o] + Q[o])))
template <int WHICH>
__device__ __forceinline__ void edge_pre2(const float (&hd)[PSI_D], f2 (&P)[PSI_D / 2]) {
    bias2(edge_mlp<WHICH>().b1, P);
    mv2<PSI_D>(edge_mlp_t<WHICH>().W1iT, hd, P);
}
template <int WHICH>
__device__ __forceinline__ void edge_q2(const float (&hs)[PSI_D], f2 (&Q)[PSI_D / 2]) {
#pragma unroll
    for (int q = 0; q < PSI_D / 2; ++q) Q[q] = pk(0.f, 0.f);
    mv2<PSI_D>(edge_mlp_t<WHICH>().W1jT, hs, Q);
}
template <int WHICH, int ATTR>
__device__ __forceinline__ void edge_z2(const f2 (&P)[PSI_D / 2], const f2 (&Q)[PSI_D / 2], const int4& rec, f2 (&z)[PSI_D / 2]) {
    const EdgeMLPT& W = edge_mlp_t<WHICH>();
    const float a[3] = {__int_as_float(rec.y), __int_as_float(rec.z), __int_as_float(rec.w)};
#pragma unroll
    for (int q = 0; q < PSI_D / 2; ++q) z[q] = fadd2(P[q], Q[q]);
    mv2<ATTR>(W.W1aT, a, z);
}
// second edge layer on the aggregated hidden sums: mp = W2·S + deg·b2
template <int WHICH>
__device__ __forceinline__ void edge_post(const float (&S)[PSI_D], int deg, float (&mp)[PSI_D]) {
    const float fdeg = (float)deg;
    f2 acc[PSI_D / 2];
    bias2(edge_mlp<WHICH>().b2, acc);
#pragma unroll
    for (int q = 0; q < PSI_D / 2; ++q) acc[q] = fmul2(pk(fdeg, fdeg), acc[q]);
    mv2<PSI_D>(edge_mlp_t<WHICH>().W2T, S, acc);
    unpack10(acc, mp);
}
// float-array forms (VJP prepare, parameter gradients)
template <int WHICH>
__device__ __forceinline__ void edge_pre(const float (&hd)[PSI_D], float (&P)[PSI_D]) {
    f2 p2[PSI_D / 2];
    edge_pre2<WHICH>(hd, p2);
    unpack10(p2, P);
}
template <int WHICH>
__device__ __forceinline__ void edge_q(const float (&hs)[PSI_D], float (&Q)[PSI_D]) {
    f2 q2[PSI_D / 2];
    edge_q2<WHICH>(hs, q2);
    unpack10(q2, Q);
}
template <int WHICH, int ATTR>
__device__ __forceinline__ void edge_z(const float (&P)[PSI_D], const float (&Q)[PSI_D], const int4& rec, float (&z)[PSI_D]) {
    f2 p2[PSI_D / 2], q2[PSI_D / 2], z2[PSI_D / 2];
    pack10(P, p2);
    pack10(Q, q2);
    edge_z2<WHICH, ATTR>(p2, q2, rec, z2);
    unpack10(z2, z);
}

// ---- pre-pass: Q[w][node] = W1j_w · h[node], rows padded to PSI_QPITCH floats (16-byte aligned rows for the gather) ------------
// One warp per 32 nodes.  The 32 input rows (1280 contiguous bytes) are read with coalesced 8-byte loads and transposed through
// shared memory; the output rows go back the same way (coalesced 16-byte stores).
template <int NQ>
__global__ void __launch_bounds__(PSI_NODE_BLOCK) k_layer_pre(int N, const float* __restrict__ h, float* __restrict__ Q, const int* __restrict__ done) {
    __shared__ __align__(16) float stage[(PSI_NODE_BLOCK / 32) * PSI_STAGE_FLOATS];
    if (done != nullptr && *done) return;
    const int lane = threadIdx.x & 31;
    const int node0 = blockIdx.x * PSI_NODE_BLOCK + (threadIdx.x & ~31);
    if (node0 >= N) return;                                        // warp-uniform
    const int rows = min(32, N - node0);
    float* st = stage + (threadIdx.x >> 5) * PSI_STAGE_FLOATS;
    const float2* src = reinterpret_cast<const float2*>(h + (int64_t)node0 * PSI_D);
#pragma unroll
    for (int p = 0; p < 5; ++p) {
        const int k = p * 32 + lane;
        if (k < rows * 5) reinterpret_cast<float2*>(st)[k] = __ldg(src + k);
    }
    __syncwarp();
    float hi[PSI_D];
#pragma unroll
    for (int q = 0; q < PSI_D / 2; ++q) {
        const float2 t = reinterpret_cast<const float2*>(st)[lane * 5 + q];
        hi[2 * q] = t.x; hi[2 * q + 1] = t.y;
    }
    __syncwarp();
#pragma unroll
    for (int w = 0; w < NQ; ++w) {
        f2 q2[PSI_D / 2];
        if (w == 0) edge_q2<0>(hi, q2);
        else if (w == 1) edge_q2<1>(hi, q2);
        else edge_q2<2>(hi, q2);
        float q[PSI_D];
        unpack10(q2, q);
        float4* row = reinterpret_cast<float4*>(st + lane * PSI_QPITCH);
        row[0] = make_float4(q[0], q[1], q[2], q[3]);
        row[1] = make_float4(q[4], q[5], q[6], q[7]);
        row[2] = make_float4(q[8], q[9], 0.f, 0.f);
        __syncwarp();
        float4* dst = reinterpret_cast<float4*>(Q + ((int64_t)w * N + node0) * PSI_QPITCH);
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const int k = p * 32 + lane;
            if (k < rows * 3) dst[k] = reinterpret_cast<const float4*>(st)[k];
        }
        __syncwarp();
    }
}

// ---- cooperative row gather ---------------------------------------------------------------------------------------
// A trip needs 32 rows of PSI_QPITCH = 12 floats (one per lane).  They are fetched as 96 16-byte pieces: lane l fetches pieces
// l + 32p (p = 0..2); piece k belongs to row-slot k / 3 (the lane that wants it), quarter k % 3.  Consecutive lanes read consecutive
// pieces of consecutive rows: a handful of cache lines per load instruction instead of 32.  The pieces are staged in shared memory
// in order (conflict-free 16-byte stores); lane l reads its row back at float offset 12·l (conflict-free as well).
// Latency: rows are requested TWO trips ahead of their use (two rotating register sets), edge records three trips ahead (four
// sets) — with rows one trip ahead the kernel sat on the long scoreboard 53 % of the time at 18 resident warps per SM (ncu,
// profiles/r02_a_operator.md); a cp.async (LDGSTS) ring instead of registers doubled the shared-memory wavefronts and was slower.
struct CoopMap {
    int r[3];      // lane whose row piece p of this lane belongs to: (32p + lane) / 3
    int c[3];      // quarter of the row: (32p + lane) % 3
};
__device__ __forceinline__ void coop_map(int lane, CoopMap& M) {
#pragma unroll
    for (int p = 0; p < 3; ++p) { M.r[p] = (p * 32 + lane) / 3; M.c[p] = (p * 32 + lane) - 3 * M.r[p]; }
}
struct __align__(16) WarpStage {
    float rows[PSI_STAGE_FLOATS];
};
// pub: the row (in units of PSI_QPITCH floats from `src`) this lane wants for the trip, clamped to a valid row (≥ 0)
__device__ __forceinline__ void coop_issue(const float* __restrict__ src, int pub, const CoopMap& M, float4 (&v)[3]) {
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const unsigned rr = (unsigned)__shfl_sync(0xffffffffu, pub, M.r[p]);
        v[p] = __ldg(reinterpret_cast<const float4*>(src + (size_t)rr * PSI_QPITCH) + M.c[p]);
    }
}
__device__ __forceinline__ void coop_store(float* st, int lane, const float4 (&v)[3]) {
#pragma unroll
    for (int p = 0; p < 3; ++p) reinterpret_cast<float4*>(st)[p * 32 + lane] = v[p];
}
__device__ __forceinline__ void coop_row2(const float* st, int lane, f2 (&q)[PSI_D / 2]) {
    const float4 a = *reinterpret_cast<const float4*>(st + lane * PSI_QPITCH);
    const float4 b = *reinterpret_cast<const float4*>(st + lane * PSI_QPITCH + 4);
    const float2 c = *reinterpret_cast<const float2*>(st + lane * PSI_QPITCH + 8);
    q[0] = pk(a.x, a.y); q[1] = pk(a.z, a.w); q[2] = pk(b.x, b.y); q[3] = pk(b.z, b.w); q[4] = pk(c.x, c.y);
}
template <class REC> __device__ __forceinline__ REC ld_rec(const REC* p);
template <> __device__ __forceinline__ int4 ld_rec<int4>(const int4* p) { return __ldg(p); }
template <> __device__ __forceinline__ int2 ld_rec<int2>(const int2* p) { return __ldg(p); }

// Walks the slice column of every lane of the warp over one SELL list.  WARP-UNIFORM: all 32 lanes must call it (the trip count is
// the slice width).  REC is the record type (int4 message records / int2 {neighbour, mask} pairs); key(rec) → row index into `src`
// (≥ 0; lanes that do not want a row return 0); body(rec, row) is called UNCONDITIONALLY for every trip — also for the padding
// records of short columns — so the caller's arithmetic must make those a no-op (the message kernels rely on relu(NaN) =
// fmaxf(NaN, 0) = 0 with the padding attributes 0xFFFFFFFF = NaN; see aggregate()).  Keeping the body free of divergent control
// flow saves the reconvergence barriers and a register copy of every accumulator per trip.  The loop is unrolled by four so that
// the rotation of the register sets costs no moves.
template <class REC, class Key, class Body>
__device__ __forceinline__ void walk_ring(const REC* __restrict__ recs, int width, const float* __restrict__ src, int lane, WarpStage& W,
                                          const CoopMap& M, Key&& key, Body&& body) {
    if (width == 0) return;
    const REC* p = recs + lane;
    float* st = W.rows;
    REC R[4];
    float4 v[2][3];
#pragma unroll
    for (int m = 0; m < 3; ++m) R[m] = ld_rec(p + 32 * min(m, width - 1));       // clamped: a duplicate of the last record is never consumed
    R[3] = R[2];
    coop_issue(src, key(R[0]), M, v[0]);
    coop_issue(src, key(R[1]), M, v[1]);
    for (int t0 = 0; t0 < width; t0 += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + u;
            if (t < width) {                                   // warp-uniform
                coop_store(st, lane, v[u % 2]);                // rows of trip t (requested at trip t − 2)
                __syncwarp();
                R[(u + 3) % 4] = ld_rec(p + 32 * min(t + 3, width - 1));
                if (t + 2 < width) coop_issue(src, key(R[(u + 2) % 4]), M, v[u % 2]);
                f2 q[PSI_D / 2];
                coop_row2(st, lane, q);
                __syncwarp();
                body(R[u], q);
            }
        }
    }
}

// The same walk with every lane gathering its own 48-byte row (three 16-byte loads, no staging, no shuffles): more cache lines per
// load instruction (one per lane) but a third of the instructions around the loads and far fewer live registers.  Two trips per
// iteration so that both records and both rows are in flight before the arithmetic of either starts.
__device__ __forceinline__ void row_direct(const float* __restrict__ src, int rowidx, f2 (&q)[PSI_D / 2]) {
    const float4* p = reinterpret_cast<const float4*>(src + (size_t)(unsigned)rowidx * PSI_QPITCH);
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    q[0] = pk(a.x, a.y); q[1] = pk(a.z, a.w); q[2] = pk(b.x, b.y); q[3] = pk(b.z, b.w); q[4] = pk(c.x, c.y);
}
template <class REC, class Key, class Body>
__device__ __forceinline__ void walk_direct(const REC* __restrict__ recs, int width, const float* __restrict__ src, int lane, Key&& key, Body&& body) {
    const REC* p = recs + lane;
    int t = 0;
    for (; t + 1 < width; t += 2) {
        const REC r0 = ld_rec(p + 32 * t), r1 = ld_rec(p + 32 * (t + 1));
        f2 q0[PSI_D / 2], q1[PSI_D / 2];
        row_direct(src, key(r0), q0);
        row_direct(src, key(r1), q1);
        body(r0, q0);
        body(r1, q1);
    }
    if (t < width) {
        const REC r0 = ld_rec(p + 32 * t);
        f2 q0[PSI_D / 2];
        row_direct(src, key(r0), q0);
        body(r0, q0);
    }
}

// Measured on B200 (ncu gpu__time_duration, profiles/r02_b_operator_ab.md): k_layer_forward<dirichlet> at C5 (1 M nodes) 106 µs
// direct vs 148 µs cooperative (196 µs for the round-1 kernel: one thread gathering its row of h with five 8-byte loads and doing
// the full 23→10 contraction per edge), at C3 (131 k nodes) 19.7 µs vs 30 µs (30.3 µs).  The cooperative walk halves the L1
// wavefronts per trip but spends ≈ 45 more instructions per trip on shuffles, staging and synchronisation — with the packed FMAs the
// kernel is issue-bound, so the simpler walk wins.  walk_ring stays as the documented alternative (-DPSI_GATHER_DIRECT=0).
#ifndef PSI_GATHER_DIRECT
#define PSI_GATHER_DIRECT 1              // 1: every lane gathers its own row (walk_direct); 0: cooperative staged gather (walk_ring)
#endif
template <class RowIdx, class Body>
__device__ __forceinline__ void walk_list(const SellDev& L, const float* __restrict__ src, int slice, int lane, WarpStage& W, const CoopMap& M,
                                          RowIdx&& rowidx_of, Body&& body) {
    const int64_t base = L.slice_off[slice];
    const int width = (int)((L.slice_off[slice + 1] - base) >> 5);
#if PSI_GATHER_DIRECT
    walk_direct(L.recs + base, width, src, lane, [&](const int4& r) { return r.x >= 0 ? rowidx_of(r.x) : 0; }, body);
#else
    walk_ring(L.recs + base, width, src, lane, W, M, [&](const int4& r) { return r.x >= 0 ? rowidx_of(r.x) : 0; }, body);
#endif
}

// Sub-wave grids (a few hundred CTAs: C0/C1/C2-size batches): the separate pre-pass launch costs more (≈ 4–6 µs of launch + latency)
// than recomputing Q_j = W1j·h_j per edge in a kernel that has one warp per scheduler and nothing to hide latency with.  Same record
// walk, but every lane gathers the neighbour's 40-byte row of h itself and forms Q with the canonical chain edge_q2 — the values, and
// therefore the results, are bit-identical to the pre-pass path.  `want` = false: the lane does not aggregate over this list (row 0).
template <int WHICH, class Body>
__device__ __forceinline__ void walk_direct_h(const int4* __restrict__ recs, int width, const float* __restrict__ h, int lane, bool want, Body&& body) {
    const int4* p = recs + lane;
    int t = 0;
    for (; t + 1 < width; t += 2) {
        const int4 r0 = __ldg(p + 32 * t), r1 = __ldg(p + 32 * (t + 1));
        float h0[PSI_D], h1[PSI_D];
        load_row(h, (want && r0.x >= 0) ? r0.x : 0, h0);
        load_row(h, (want && r1.x >= 0) ? r1.x : 0, h1);
        f2 q0[PSI_D / 2], q1[PSI_D / 2];
        edge_q2<WHICH>(h0, q0);
        edge_q2<WHICH>(h1, q1);
        body(r0, q0);
        body(r1, q1);
    }
    if (t < width) {
        const int4 r0 = __ldg(p + 32 * t);
        float h0[PSI_D];
        load_row(h, (want && r0.x >= 0) ? r0.x : 0, h0);
        f2 q0[PSI_D / 2];
        edge_q2<WHICH>(h0, q0);
        body(r0, q0);
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
    f2 acc[PSI_D / 2];
    bias2(cW.up_b1, acc);
    mv2<PSI_D>(cWT.up_W1T, hi, acc);
    mv2<PSI_D>(cWT.up_W1T + PSI_D, mT, acc);
    mv2<PSI_D>(cWT.up_W1T + 2 * PSI_D, mF, acc);
    mv2<PRB>(cWT.up_W1T + 3 * PSI_D, prb, acc);
    float t[PSI_D];
    unpack10(acc, t);
    hmask = 0;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        if (t[o] > 0.f) hmask |= (1u << o);
        hid[o] = fmaxf(t[o], 0.f);
    }
    bias2(cW.up_b2, acc);
    mv2<PSI_D>(cWT.up_W2T, hid, acc);
    unpack10(acc, m);
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
    f2 acc[PSI_D / 2];
    bias2(cW.un_b1, acc);
    mv2<PSI_D>(cWT.un_W1T, hi, acc);
    mv2<PSI_D>(cWT.un_W1T + PSI_D, mN, acc);
    mv2<3>(cWT.un_W1T + 2 * PSI_D, prb, acc);
    mv2<2>(cWT.un_W1T + 2 * PSI_D + 3, nv, acc);
    float t[PSI_D];
    unpack10(acc, t);
    hmask = 0;
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) {
        if (t[o] > 0.f) hmask |= (1u << o);
        hid[o] = fmaxf(t[o], 0.f);
    }
    bias2(cW.un_b2, acc);
    mv2<PSI_D>(cWT.un_W2T, hid, acc);
    unpack10(acc, m);
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
// Lanes that do not aggregate over a list (Dirichlet rows, Neumann rows over list T, rows beyond the produced range) run the same
// instructions on row 0 and discard the result.  Padding records contribute relu(NaN) = 0 to S and nothing to deg.
template <int KIND, bool NOPRE = false>
__device__ __forceinline__ void aggregate(const GraphDev& G, const float* __restrict__ h, const float* __restrict__ Q, int node, int cls,
                                          WarpStage& st, const CoopMap& M, float (&mT)[PSI_D], float (&mF)[PSI_D]) {
    constexpr int ATTR = KindTraits<KIND>::ATTR;
    constexpr bool NEU = KindTraits<KIND>::has_neumann;
    const int lane = threadIdx.x & 31, slice = node >> 5;
    const int N = G.N;
    f2 P[PSI_D / 2], S[PSI_D / 2];
    float Sf[PSI_D], hi[PSI_D];
    int deg = 0;
    // S += relu(z): the ReLU has no packed form (two FMNMX), the accumulation has
    auto relu_acc = [&](const int4& rec, const f2 (&z)[PSI_D / 2]) {
#pragma unroll
        for (int q = 0; q < PSI_D / 2; ++q) {
            float a, b;
            upk(z[q], a, b);
            S[q] = fadd2(S[q], pk(fmaxf(a, 0.f), fmaxf(b, 0.f)));
        }
        deg += (rec.x >= 0) ? 1 : 0;
    };
    // ---- list T: messages Φ→ into interior destinations (neighbour = row index of the entry) ----
#pragma unroll
    for (int q = 0; q < PSI_D / 2; ++q) { S[q] = pk(0.f, 0.f); P[q] = pk(0.f, 0.f); }
    if (cls == 0) {
        load_row(h, node, hi);                   // re-read per list instead of held across the walks (register pressure)
        edge_pre2<0>(hi, P);
    }
    auto body_T = [&](const int4& rec, const f2 (&q)[PSI_D / 2]) {
        f2 z[PSI_D / 2];
        edge_z2<0, ATTR>(P, q, rec, z);
        relu_acc(rec, z);
    };
    if (NOPRE) {
        const int64_t base = G.T.slice_off[slice];
        walk_direct_h<0>(G.T.recs + base, (int)((G.T.slice_off[slice + 1] - base) >> 5), h, lane, cls == 0, body_T);
    } else {
        walk_list(G.T, Q, slice, lane, st, M, [&](int j) { return cls == 0 ? j : 0; }, body_T);
    }
    unpack10(S, Sf);
    if (cls == 0) edge_post<0>(Sf, deg, mT);
    asm volatile("" ::: "memory");               // keep the compiler from carrying the first read of the row across the walk
    // ---- list F: messages Φ← into interior destinations, Φ_neumann into Neumann destinations ----
#pragma unroll
    for (int q = 0; q < PSI_D / 2; ++q) S[q] = pk(0.f, 0.f);
    deg = 0;
    if (cls == 0) {
        load_row(h, node, hi);
        edge_pre2<1>(hi, P);
    }
    if (NEU && cls == 2) {
        load_row(h, node, hi);
        edge_pre2<2>(hi, P);
    }
    if (NEU) {
        // the Neumann rows of a warp use another edge MLP: W1a differs per lane class, so this walk keeps a (divergent) branch
        walk_list(G.F, Q, slice, lane, st, M,
                  [&](int j) { return cls == 0 ? N + j : (cls == 2 ? 2 * N + j : 0); },
                  [&](const int4& rec, const f2 (&q)[PSI_D / 2]) {
                      f2 z[PSI_D / 2];
                      if (cls == 2) edge_z2<2, ATTR>(P, q, rec, z);
                      else edge_z2<1, ATTR>(P, q, rec, z);
                      relu_acc(rec, z);
                  });
    } else {
        auto body_F = [&](const int4& rec, const f2 (&q)[PSI_D / 2]) {
            f2 z[PSI_D / 2];
            edge_z2<1, ATTR>(P, q, rec, z);
            relu_acc(rec, z);
        };
        if (NOPRE) {
            const int64_t base = G.F.slice_off[slice];
            walk_direct_h<1>(G.F.recs + base, (int)((G.F.slice_off[slice + 1] - base) >> 5), h, lane, cls == 0, body_F);
        } else {
            walk_list(G.F, Q, slice, lane, st, M, [&](int j) { return cls == 0 ? N + j : 0; }, body_F);
        }
    }
    unpack10(S, Sf);
    if (cls == 0) edge_post<1>(Sf, deg, mF);
    if (NEU && cls == 2) edge_post<2>(Sf, deg, mF);
    asm volatile("" ::: "memory");
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
    for (int i = 0; i < PSI_D; ++i) c[i] = rk[i] * hi[i];                       // cat[reset*H, to, from, prb]
    f2 ac[PSI_D / 2];
    bias2(cW.gc_b, ac);
    mv2<30 + PRB>(cWT.gcT, c, ac);
    float a[PSI_D];
    unpack10(ac, a);
#pragma unroll
    for (int o = 0; o < PSI_D; ++o) out[o] = fmaf(zk[o], tanhf(a[o]), hi[o]);   // H + alpha*corr
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

template <int KIND, bool EPI, bool NOPRE = false>
__global__ void __launch_bounds__(PSI_NODE_BLOCK, PSI_OP_MIN_CTAS)
k_layer_forward(GraphDev G, const float* __restrict__ h, const float* __restrict__ h0, const float* __restrict__ Q, float* __restrict__ out,
                SolverEpi E) {
    __shared__ WarpStage stage[PSI_NODE_BLOCK / 32];
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
        CoopMap M;
        coop_map(lane, M);
        float mT[PSI_D], mF[PSI_D];
#pragma unroll
        for (int o = 0; o < PSI_D; ++o) { mT[o] = 0.f; mF[o] = 0.f; }
        static_assert(!(NOPRE && KindTraits<KIND>::has_neumann), "the fused per-edge product exists for the kinds without a Neumann edge MLP");
        aggregate<KIND, NOPRE>(G, h, Q, node, cls, stage[threadIdx.x >> 5], M, mT, mF);
        if (valid) {
            load_row(h, node, hi);
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

// flux form of the DSS residual (dirichlet/dss/model.py:137-145): out_i = Σ_{e = (i → j)} a_e (v_j − v_i) in CSR order (the reference
// scatter_adds the per-edge fluxes: float atomics on CUDA).  TRANSPOSE: the adjoint out_j = Σ_{e = (i → j)} a_e v_i − v_j Σ_{e = (j → ·)} a_e.
template <bool TRANSPOSE>
__global__ void __launch_bounds__(PSI_NODE_BLOCK) k_flux(GraphDev G, const float* __restrict__ v, float* __restrict__ out) {
    const int node = blockIdx.x * PSI_NODE_BLOCK + threadIdx.x;
    if (node >= G.N) return;
    const float vi = __ldg(v + node);
    float s = 0.f, t = 0.f;
    {
        const SellDev& L = G.Ar;
        const int64_t base = L.slice_off[node >> 5];
        const int width = (int)((L.slice_off[(node >> 5) + 1] - base) >> 5);
        const int2* p = L.recs2 + base + (node & 31);
        for (int k = 0; k < width; ++k) {
            const int2 rec = __ldg(p + (int64_t)k * 32);
            if (rec.x < 0) continue;
            if (TRANSPOSE) s += __int_as_float(rec.y);
            else s = fmaf(__int_as_float(rec.y), __ldg(v + rec.x) - vi, s);
        }
    }
    if (TRANSPOSE) {
        const SellDev& L = G.Ac;
        const int64_t base = L.slice_off[node >> 5];
        const int width = (int)((L.slice_off[(node >> 5) + 1] - base) >> 5);
        const int2* p = L.recs2 + base + (node & 31);
        for (int k = 0; k < width; ++k) {
            const int2 rec = __ldg(p + (int64_t)k * 32);
            if (rec.x >= 0) t = fmaf(__int_as_float(rec.y), __ldg(v + rec.x), t);
        }
        out[node] = fmaf(-s, vi, t);
    } else {
        out[node] = s;
    }
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
