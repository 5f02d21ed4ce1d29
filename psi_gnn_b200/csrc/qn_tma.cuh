// qn_tma.cuh — the two history-streaming kernels of the Broyden update as TMA-fed persistent kernels.
//
// The U/V history (2·n vectors of N·d floats) is read twice per step and dominates HBM traffic.  Both
// passes are pure streams with no reuse, so the job is to keep enough bytes in flight per SM independent of register
// pressure: one persistent CTA per SM, a producer warp issuing 1-D bulk copies (cp.async.bulk → UBLKCP) of U_k / V_k tiles
// into a shared-memory ring, full/empty mbarriers per stage, eight consumer warps doing the FMAs out of shared memory.
//   pass 1  k_qn_dots_tma : a = Uᵀδx, c = Vᵀδg, e = Vᵀg        work item = (4096-element chunk, range of 32 history vectors)
//   pass 2  k_qn_axpy_tma : v = −δx + V·a, w = U·c, t = U·e, then u_n, the new update and x ← x + update in the same sweep;
//                           each CTA owns an equal contiguous element range (±16 B)
// Summation orders are fixed (xor-shuffle tree, then warps in order, then chunks in order): results are deterministic.
#pragma once
#include "common.cuh"
#include "broyden.cuh"

#define TMA_CONSUMERS 256                    // 8 consumer warps
#define TMA_THREADS (TMA_CONSUMERS + 32)     // + 1 producer warp
#define DOTS_CH 4096                         // elements per chunk (16 KB per vector), four float4 per consumer thread
#define DOTS_STAGES 6                        // 6 × (16 KB U + 16 KB V) = 192 KB
#define DOTS_KR 32                           // history vectors per work item (upper bound; the launch picks 8, 16 or 32)
#ifndef DOTS_PRODUCER_LANES
#define DOTS_PRODUCER_LANES 4                // producer lanes issuing fills in parallel (< DOTS_STAGES)
#endif
#define DOTS_KB 8                            // ks per cross-warp reduction batch
#define AXPY_TILE 2048                       // max elements per tile (8 KB per vector), two float4 per consumer thread
#define AXPY_STAGES 11                       // 11 × (8 KB U + 8 KB V) = 176 KB
#ifndef AXPY_PRODUCER_LANES
#define AXPY_PRODUCER_LANES 8                // producer lanes issuing fills in parallel (< AXPY_STAGES)
#endif

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// 1-D bulk copy global → shared, completion signalled on an mbarrier (bytes multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(TMA_CONSUMERS) : "memory"); }

struct TmaRing {
    uint64_t* full;
    uint64_t* empty;
};

// ---- pass 1 ---------------------------------------------------------------------------------------------------------
// partial[(k*3+q)*num_chunks + chunk], num_chunks = stride / DOTS_CH
__global__ void __launch_bounds__(TMA_THREADS, 1)
k_qn_dots_tma(QnHistory H, int nhist, const float* __restrict__ dx, const float* __restrict__ dg, const float* __restrict__ g,
              float* __restrict__ partial, int num_chunks, const int* __restrict__ done, int kr) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);                                   // [STAGES][2][DOTS_CH]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)DOTS_STAGES * 2 * DOTS_CH * sizeof(float));
    uint64_t* empty = full + DOTS_STAGES;
    float* red = reinterpret_cast<float*>(empty + DOTS_STAGES);                         // [2][DOTS_KB][3][8]
    float* red_x = red + 2 * DOTS_KB * 3 * 8;                                           // [2][2][8]: ⟨δx,δg⟩, ⟨δx,g⟩ of a chunk
    if (*done) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < DOTS_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], TMA_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // kr history vectors per work item: small problems (few chunks) get finer items so that every SM has work — each (vector, chunk)
    // dot product is formed inside one item whatever kr is, so the results do not depend on it
    const int kranges = max(1, (nhist + kr - 1) / kr);                 // at least one item per chunk: it also carries ⟨δx,δg⟩, ⟨δx,g⟩
    const int items = num_chunks * kranges;
    if (warp == TMA_CONSUMERS / 32) {
        // ---- producer ----
        // DOTS_PRODUCER_LANES lanes (fewer than stages) take consecutive fills, so that their wait → expect → copy chains overlap.
        // Rounds are warp-synchronous: every fill of a round is issued before the next round starts, hence when a lane waits for the
        // slot of fill f, fill f − STAGES has been issued and the parity it waits on is that of the slot's CURRENT phase (a lane running
        // two phases ahead would see a stale parity as "free").
        uint32_t fill0 = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int c = item % num_chunks, r = item / num_chunks;
            const int k0 = r * kr, k1 = min(nhist, k0 + kr);
            const int64_t e0 = (int64_t)c * DOTS_CH;
            for (int kk = k0; kk < k1; kk += DOTS_PRODUCER_LANES) {
                const int k = kk + lane;
                if (lane < DOTS_PRODUCER_LANES && k < k1) {
                    const uint32_t fill = fill0 + (uint32_t)(k - k0);
                    const uint32_t s = fill % DOTS_STAGES, use = fill / DOTS_STAGES;
                    if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
                    mbar_expect_tx(&full[s], 2 * DOTS_CH * sizeof(float));
                    float* dst = ring + (size_t)s * 2 * DOTS_CH;
                    bulk_g2s(dst, hist_u(H, k) + e0, DOTS_CH * sizeof(float), &full[s]);
                    bulk_g2s(dst + DOTS_CH, hist_v(H, k) + e0, DOTS_CH * sizeof(float), &full[s]);
                }
                __syncwarp();
            }
            fill0 += (uint32_t)(k1 - k0);
        }
        return;
    }
    // ---- consumers ----
    uint32_t fill = 0;
    int buf = 0, xbuf = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int c = item % num_chunks, r = item / num_chunks;
        const int k0 = r * kr, k1 = min(nhist, k0 + kr);
        const int64_t e0 = (int64_t)c * DOTS_CH;
        const float4* pdx = reinterpret_cast<const float4*>(dx + e0);
        const float4* pdg = reinterpret_cast<const float4*>(dg + e0);
        const float4* pg = reinterpret_cast<const float4*>(g + e0);
        constexpr int R = DOTS_CH / 4 / TMA_CONSUMERS;          // float4 per thread per vector
        float4 rdx[R], rdg[R], rg[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            rdx[j] = __ldg(pdx + tid + j * TMA_CONSUMERS);
            rdg[j] = __ldg(pdg + tid + j * TMA_CONSUMERS);
            rg[j] = __ldg(pg + tid + j * TMA_CONSUMERS);
        }
        if (r == 0) {
            // rows 3·nhist and 3·nhist+1: ⟨δx,δg⟩ and ⟨δx,g_n⟩.  With them s = ⟨v_n,δg⟩ and p = ⟨v_n,g_n⟩ follow from this pass by
            // linearity (v_n = −δx + Σ a_k V_k ⇒ s = −⟨δx,δg⟩ + Σ a_k c_k, p = −⟨δx,g⟩ + Σ a_k e_k): pass 2 needs no reduction of its own.
            float d1 = 0.f, d2 = 0.f;
#pragma unroll
            for (int j = 0; j < R; ++j) { d1 = dot4(rdx[j], rdg[j], d1); d2 = dot4(rdx[j], rg[j], d2); }
            d1 = warp_sum(d1); d2 = warp_sum(d2);
            if (lane == 0) { red_x[(xbuf * 2 + 0) * 8 + warp] = d1; red_x[(xbuf * 2 + 1) * 8 + warp] = d2; }
            consumer_bar();
            if (tid < 2) {
                const float* rp = red_x + (xbuf * 2 + tid) * 8;
                float sum = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) sum += rp[w];
                partial[((int64_t)3 * nhist + tid) * num_chunks + c] = sum;
            }
            xbuf ^= 1;
        }
        for (int kb0 = k0; kb0 < k1; kb0 += DOTS_KB) {
            const int nb = min(DOTS_KB, k1 - kb0);
            for (int kk = 0; kk < nb; ++kk, ++fill) {
                const uint32_t s = fill % DOTS_STAGES, use = fill / DOTS_STAGES;
                mbar_wait(&full[s], use & 1);
                const float4* su = reinterpret_cast<const float4*>(ring + (size_t)s * 2 * DOTS_CH);
                const float4* sv = su + DOTS_CH / 4;
                float4 u[R], v[R];
#pragma unroll
                for (int j = 0; j < R; ++j) { u[j] = su[tid + j * TMA_CONSUMERS]; v[j] = sv[tid + j * TMA_CONSUMERS]; }
#ifdef PSI_EARLY_RELEASE          // diagnosis only: the round-1 order (release right behind the loads)
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
#endif
                float a = 0.f, cc = 0.f, e = 0.f;
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    a = dot4(u[j], rdx[j], a);
                    cc = dot4(v[j], rdg[j], cc);
                    e = dot4(v[j], rg[j], e);
                }
                a = warp_sum(a); cc = warp_sum(cc); e = warp_sum(e);
                // Release the slot only now: the shuffles above consumed every lane's loaded values, so all shared-memory reads of the
                // warp have completed.  (An arrive issued right behind the loads runs in another pipe and can overtake loads still in
                // flight; the refill — fastest when the history sits in L2 — would then overwrite data not yet read.)
#ifndef PSI_EARLY_RELEASE
                if (lane == 0) mbar_arrive(&empty[s]);
#endif
                if (lane == 0) {
                    float* rp = red + ((buf * DOTS_KB + kk) * 3) * 8 + warp;
                    rp[0] = a; rp[8] = cc; rp[16] = e;
                }
            }
            consumer_bar();
            if (tid < nb * 3) {
                const int kk = tid / 3, q = tid % 3;
                const float* rp = red + ((buf * DOTS_KB + kk) * 3 + q) * 8;
                float sum = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) sum += rp[w];
                partial[((int64_t)(kb0 + kk) * 3 + q) * num_chunks + c] = sum;
            }
            buf ^= 1;
        }
    }
}

// ---- pass 2 ---------------------------------------------------------------------------------------------------------
// v_n = −δx + V·a ; w = −δg + U·c ; t = U·e ; u_n = (δx − w)/s ; update = g_n − t − u_n·p ; x ← x + update ; δx ← x_new − x.
// s = ⟨v_n,δg⟩ and p = ⟨v_n,g_n⟩ come from the fp64 row sums of pass 1 (dbuf), so the whole rank-one update and the step are one
// sweep with no trailing reduction.  Each CTA owns float4 range [q0, q1) of the vectors, split into equal tiles of ≤ AXPY_TILE floats.
__global__ void __launch_bounds__(TMA_THREADS, 1)
k_qn_axpy_tma(QnHistory H, int nhist, int n, const float* __restrict__ coef, int cap, const double* __restrict__ dbuf, float* __restrict__ dx_io,
              const float* __restrict__ dg, const float* __restrict__ g, float* __restrict__ x, float* __restrict__ best_x,
              float* __restrict__ xtrace_next, int64_t total4, QnCtrl* __restrict__ ctrl) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);                                   // [STAGES][2][AXPY_TILE]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)AXPY_STAGES * 2 * AXPY_TILE * sizeof(float));
    uint64_t* empty = full + AXPY_STAGES;
    double* s_red = reinterpret_cast<double*>(empty + AXPY_STAGES);                     // [2][9]
    float* s_coef = reinterpret_cast<float*>(s_red + 18);                               // [3][nhist]
    const int improved = ctrl->improved, done = ctrl->done;
    if (done && !improved) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t per = (total4 + gridDim.x - 1) / gridDim.x;
    const int64_t q0 = min(total4, (int64_t)blockIdx.x * per), q1 = min(total4, q0 + per);
    if (improved) {                        // lowest_xest = x_est.clone()  (solver.py:172)
        for (int64_t q = q0 + tid; q < q1; q += TMA_THREADS)
            reinterpret_cast<float4*>(best_x)[q] = reinterpret_cast<const float4*>(x)[q];
    }
    if (done) return;
    if (tid == 0) {
        for (int s = 0; s < AXPY_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], TMA_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < nhist; i += TMA_THREADS) {
        s_coef[i] = coef[i];
        s_coef[nhist + i] = coef[cap + i];
        s_coef[2 * nhist + i] = coef[2 * cap + i];
    }
    {   // s = −⟨δx,δg⟩ + Σ_k a_k c_k ,  p = −⟨δx,g⟩ + Σ_k a_k e_k  in fp64, identical order in every CTA (and on every rank)
        double ps = 0.0, pp = 0.0;
        for (int k = tid; k < nhist; k += TMA_THREADS) {
            const double ak = dbuf[3 * k];
            ps = fma(ak, dbuf[3 * k + 1], ps);
            pp = fma(ak, dbuf[3 * k + 2], pp);
        }
        ps = warp_sum_d(ps);
        pp = warp_sum_d(pp);
        if (lane == 0) { s_red[warp] = ps; s_red[9 + warp] = pp; }
    }
    __syncthreads();
    double sd = -dbuf[3 * nhist], pd = -dbuf[3 * nhist + 1];
#pragma unroll
    for (int w = 0; w < TMA_THREADS / 32; ++w) { sd += s_red[w]; pd += s_red[9 + w]; }
    const float s = (float)sd, p = (float)pd;            // torch.einsum(...) / matvec produce fp32 scalars (solver.py:187,192)
    if (blockIdx.x == 0 && tid == 0) { ctrl->s = sd; ctrl->p = pd; }
    const int64_t len4 = q1 - q0;
    const int ntiles = (int)((len4 * 4 + AXPY_TILE - 1) / AXPY_TILE);
    const int tile4 = ntiles > 0 ? (int)((len4 + ntiles - 1) / ntiles) : 0;            // ≤ AXPY_TILE/4 float4 per tile
    if (warp == TMA_CONSUMERS / 32) {
        // Producer: AXPY_PRODUCER_LANES lanes take consecutive fills (fewer lanes than stages: the fills of one round land in distinct
        // stages).  One lane's wait → expect → copy → copy chain costs ≈ 0.3 µs; with a single producer lane that chain caps the SM at
        // one fill per 0.3 µs, which is below the HBM rate once a fill is < 16 KB (small batches: 4.4 KB per vector at 16 k nodes).
        const uint32_t total = (uint32_t)ntiles * (uint32_t)nhist;
        for (uint32_t f0 = 0; f0 < total; f0 += AXPY_PRODUCER_LANES) {
            const uint32_t fill = f0 + lane;
            if (lane < AXPY_PRODUCER_LANES && fill < total) {
                const uint32_t t = fill / (uint32_t)nhist;
                const int k = nhist - 1 - (int)(fill - t * (uint32_t)nhist);   // descending: pass 1 just streamed the newest vectors
                                                                               // last, the tail of the history is what L2 still holds
                const int64_t t0 = q0 + (int64_t)t * tile4;
                const uint32_t bytes = (uint32_t)(min((int64_t)tile4, q1 - t0) * 16);
                const uint32_t st = fill % AXPY_STAGES, use = fill / AXPY_STAGES;
                if (use > 0) mbar_wait(&empty[st], (use - 1) & 1);
                mbar_expect_tx(&full[st], 2 * bytes);
                float* dst = ring + (size_t)st * 2 * AXPY_TILE;
                bulk_g2s(dst, hist_u(H, k) + t0 * 4, bytes, &full[st]);
                bulk_g2s(dst + AXPY_TILE, hist_v(H, k) + t0 * 4, bytes, &full[st]);
            }
            __syncwarp();
        }
        return;
    }
    float* vn_dst = hist_v(H, n - 1);
    float* un_dst = hist_u(H, n - 1);
    uint32_t fill = 0;
    for (int t = 0; t < ntiles; ++t) {
        const int64_t t0 = q0 + (int64_t)t * tile4;
        const int cnt4 = (int)min((int64_t)tile4, q1 - t0);
        const bool act0 = tid < cnt4, act1 = tid + TMA_CONSUMERS < cnt4;
        float4 av[2], aw[2], at[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) { av[j] = make_float4(0.f, 0.f, 0.f, 0.f); aw[j] = av[j]; at[j] = av[j]; }
        for (int k = nhist - 1; k >= 0; --k, ++fill) {
            const uint32_t st = fill % AXPY_STAGES, use = fill / AXPY_STAGES;
            mbar_wait(&full[st], use & 1);
            const float4* su = reinterpret_cast<const float4*>(ring + (size_t)st * 2 * AXPY_TILE);
            const float a = s_coef[k], c = s_coef[nhist + k], e = s_coef[2 * nhist + k];
            float4 u[2], v[2];
            // rows beyond cnt4 hold stale ring data: they are read (harmless) but never stored
            u[0] = su[tid]; u[1] = su[tid + TMA_CONSUMERS];
            v[0] = su[AXPY_TILE / 4 + tid]; v[1] = su[AXPY_TILE / 4 + tid + TMA_CONSUMERS];
#ifdef PSI_EARLY_RELEASE
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
#endif
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                av[j].x = fmaf(a, v[j].x, av[j].x); av[j].y = fmaf(a, v[j].y, av[j].y); av[j].z = fmaf(a, v[j].z, av[j].z); av[j].w = fmaf(a, v[j].w, av[j].w);
                aw[j].x = fmaf(c, u[j].x, aw[j].x); aw[j].y = fmaf(c, u[j].y, aw[j].y); aw[j].z = fmaf(c, u[j].z, aw[j].z); aw[j].w = fmaf(c, u[j].w, aw[j].w);
                at[j].x = fmaf(e, u[j].x, at[j].x); at[j].y = fmaf(e, u[j].y, at[j].y); at[j].z = fmaf(e, u[j].z, at[j].z); at[j].w = fmaf(e, u[j].w, at[j].w);
            }
            // release the slot after the FMAs have consumed the loaded values (see pass 1): no shared-memory read of the warp is in flight
#ifndef PSI_EARLY_RELEASE
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
#endif
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (j == 0 ? act0 : act1) {
                const int64_t q = t0 + tid + j * TMA_CONSUMERS;
                const float4 vdx = reinterpret_cast<const float4*>(dx_io)[q];
                const float4 vdg = reinterpret_cast<const float4*>(dg)[q];
                const float4 vg = reinterpret_cast<const float4*>(g)[q];
                const float4 xv = reinterpret_cast<const float4*>(x)[q];
                // vT = −δx + Σ a_k V_k  (rmatvec, solver.py:104) ; vT[vT != vT] = 0  (solver.py:188)
                float4 vn = make_float4(-vdx.x + av[j].x, -vdx.y + av[j].y, -vdx.z + av[j].z, -vdx.w + av[j].w);
                vn.x = (vn.x != vn.x) ? 0.f : vn.x; vn.y = (vn.y != vn.y) ? 0.f : vn.y;
                vn.z = (vn.z != vn.z) ? 0.f : vn.z; vn.w = (vn.w != vn.w) ? 0.f : vn.w;
                // u = (δx − matvec(δg)) / ⟨vT,δg⟩ = (δx − (−δg + Σ c_k U_k)) / s ; u[u != u] = 0   (solver.py:114,187,189)
                float4 un = make_float4(vdx.x - (-vdg.x + aw[j].x), vdx.y - (-vdg.y + aw[j].y), vdx.z - (-vdg.z + aw[j].z),
                                        vdx.w - (-vdg.w + aw[j].w));
                un.x = __fdiv_rn(un.x, s); un.y = __fdiv_rn(un.y, s); un.z = __fdiv_rn(un.z, s); un.w = __fdiv_rn(un.w, s);
                un.x = (un.x != un.x) ? 0.f : un.x; un.y = (un.y != un.y) ? 0.f : un.y;
                un.z = (un.z != un.z) ? 0.f : un.z; un.w = (un.w != un.w) ? 0.f : un.w;
                // update = −matvec(U[:n], V[:n], g) = −(−g + Σ_{k<n-1} U_k e_k + u_n·p)   (solver.py:192)
                float4 upd;
                upd.x = -(-vg.x + fmaf(un.x, p, at[j].x)); upd.y = -(-vg.y + fmaf(un.y, p, at[j].y));
                upd.z = -(-vg.z + fmaf(un.z, p, at[j].z)); upd.w = -(-vg.w + fmaf(un.w, p, at[j].w));
                // line_search with s = 1: x_est = x0 + update ; delta_x = x_est − x0   (solver.py:89,94)
                const float4 xn = make_float4(xv.x + upd.x, xv.y + upd.y, xv.z + upd.z, xv.w + upd.w);
                reinterpret_cast<float4*>(vn_dst)[q] = vn;
                reinterpret_cast<float4*>(un_dst)[q] = un;
                reinterpret_cast<float4*>(x)[q] = xn;
                reinterpret_cast<float4*>(dx_io)[q] = make_float4(xn.x - xv.x, xn.y - xv.y, xn.z - xv.z, xn.w - xv.w);
                if (xtrace_next != nullptr) reinterpret_cast<float4*>(xtrace_next)[q] = xn;
            }
        }
    }
}

static inline size_t dots_tma_smem() { return (size_t)DOTS_STAGES * 2 * DOTS_CH * 4 + 2 * DOTS_STAGES * 8 + 2 * DOTS_KB * 3 * 8 * 4 + 2 * 2 * 8 * 4; }
static inline size_t axpy_tma_smem(int nhist) { return (size_t)AXPY_STAGES * 2 * AXPY_TILE * 4 + 2 * AXPY_STAGES * 8 + 18 * 8 + (size_t)3 * (nhist > 0 ? nhist : 1) * 4; }
