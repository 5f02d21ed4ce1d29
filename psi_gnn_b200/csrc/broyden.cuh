// broyden.cuh — quasi-Newton bookkeeping of the fixed-point solve as bandwidth-bound streaming kernels.
//
// Reference: `broyden` in */utilities/solver.py:116-207 with ls=False (s = 1), stop_mode 'rel', bsz = 1
// (the whole batch is one vector of numel = N·d).  Inverse-Jacobian approximation B = −I + U Vᵀ.
//
// Per step n (after the operator kernel produced g_n, δg and the norm partials):
//   pass 1  k_qn_dots_tma : a = U[:n-1]ᵀδx, c = V[:n-1]ᵀδg, e = V[:n-1]ᵀg_n      reads U,V once   (solver.py:103,113)
//           k_qn_fin1 : reduce partials; ‖g‖, rel; best/trace/stop rules on the device  (solver.py:162-183)
//   pass 2  k_qn_axpy_tma : v_n = −δx + V·a ; w = U·c ; t = U·e ; u_n = (δx − (−δg + w))/s ; update = −(−g_n + t + u_n·p) ; x ← x + update
//           with s = ⟨v_n,δg⟩ = −⟨δx,δg⟩ + Σ a_k c_k and p = ⟨v_n,g_n⟩ = −⟨δx,g_n⟩ + Σ a_k e_k from pass 1 by linearity (fp64)   reads U,V once
// = 4(n−1) history vectors of traffic per step (the reference's op order costs 6n−4) while still forming
// the reference's u_n, v_n, update_n from fresh dot products (no cached Vᵀg — that drifts).
// No host synchronisation inside the loop: the stop decision lives in a device control block and every
// kernel of later steps exits immediately once `done` is set; the host polls it once per chunk of steps.
#pragma once
#include "common.cuh"

#define QN_CHUNK 1024            // elements per CTA-chunk (256 threads × float4)
#define QN_THREADS 256
#define QN_MAX_SLABS 16
#define QN_AXPY_MAX_CTAS (PSI_NUM_SMS_B200 * 8)

struct QnCtrl {
    int done;            // solve finished (set by k_qn_fin1)
    int improved;        // this step produced the best iterate so far
    int nstep;           // steps executed so far
    int prot_break;
    int stop_reason;     // 1 rel<eps, 2 plateau, 3 protective break
    int best_step_rel;
    int best_step_abs;
    int pad_;
    double best_rel, best_abs, first_rel;
    double s, p;         // ⟨v_n,δg⟩, ⟨v_n,g_n⟩ of the current step (debug / forced-step API)
    int* host_done;      // mapped pinned host word set to 1 when the solve stops: the host reads it without synchronising the stream
};

struct QnHistory {       // U and V stored as slabs of `slab_vecs` contiguous vectors of `stride` floats
    float* U[QN_MAX_SLABS];
    float* V[QN_MAX_SLABS];
    int slab_vecs;
    int64_t stride;
};

__device__ __forceinline__ float* hist_u(const QnHistory& H, int k) {
    return H.U[k / H.slab_vecs] + (int64_t)(k % H.slab_vecs) * H.stride;
}
__device__ __forceinline__ float* hist_v(const QnHistory& H, int k) {
    return H.V[k / H.slab_vecs] + (int64_t)(k % H.slab_vecs) * H.stride;
}

__device__ __forceinline__ float4 ld4_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    return acc;
}

// pass 1 (k_qn_dots_tma) and pass 2 (k_qn_axpy_tma) live in qn_tma.cuh

// stopping rules of the reference loop (solver.py:160-183), evaluated by one warp in double precision
__device__ __forceinline__ void qn_decide(int lane, double n1, double n2, QnCtrl* __restrict__ ctrl, double* __restrict__ rel_trace,
                                          double* __restrict__ abs_trace, int step, double eps, double protect, int threshold) {
    // plateau rule needs max/min over the last 30 rel values (incl. this one)
    double rel = 0.0, absd = 0.0;
    bool finite = true;
    if (lane == 0) {
        const float na = (float)sqrt(n1), nb = (float)sqrt(n2);       // torch.norm(...) returns fp32
        finite = isfinite(na) && isfinite(nb);
        absd = (double)na;
        rel = absd / ((double)nb + 1e-9);
        rel_trace[step - 1] = rel;
        abs_trace[step - 1] = absd;
    }
    rel = __shfl_sync(0xffffffffu, rel, 0);
    __syncwarp();
    double mx = -1.0, mn = 1e300;
    if (step > 30) {
        // 30 most recent entries: indices step-30 .. step-1 ; lane 0 holds the newest which may not be visible
        // to other lanes through memory yet, so it is injected from the register.
        if (lane < 30) {
            const int idx = step - 30 + lane;
            const double v = (idx == step - 1) ? rel : rel_trace[idx];
            mx = v; mn = v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
    }
    if (lane == 0) {
        int improved = 0;
        if (step == 1) ctrl->first_rel = rel;
        if (rel < ctrl->best_rel) { ctrl->best_rel = rel; ctrl->best_step_rel = step; improved = 1; }
        if (absd < ctrl->best_abs) { ctrl->best_abs = absd; ctrl->best_step_abs = step; }
        ctrl->improved = improved;
        ctrl->nstep = step;
        int stop = 0;
        if (rel < eps) stop = 1;
        else if (rel < 3.0 * eps && step > 30 && mx / mn < 1.3) stop = 2;
        else if (rel > ctrl->first_rel * protect) { stop = 3; ctrl->prot_break = 1; }
        if (!finite && stop == 0) {
            // NaN/Inf residual: every comparison above is false in the reference as well, the loop runs on
            // to the threshold producing NaNs.  The outcome (best iterate so far) is unchanged by stopping now.
            stop = 4;
        }
        if (step >= threshold && stop == 0) stop = 5;      // while nstep < threshold
        if (stop) {
            ctrl->stop_reason = stop; ctrl->done = 1;
            if (ctrl->host_done != nullptr) { *reinterpret_cast<volatile int*>(ctrl->host_done) = 1; __threadfence_system(); }
        }
    }
}

// ---- finalize 1: reduce dot partials, norms, device-side stopping rules --------------------------------
// Each warp reduces rows (k,q) of the partial matrix in a grid-stride loop; block 0 / warp 0 additionally
// evaluates the reference's stopping logic (solver.py:160-183) in double precision.
__global__ void __launch_bounds__(256)
k_qn_fin1(int nhist, const float* __restrict__ partial, int num_chunks, float* __restrict__ coef /* [3][cap] */, int cap,
          double* __restrict__ dbuf /* [3·nhist + 2] fp64 row sums for pass 2 */, const float* __restrict__ norm_part, int norm_blocks,
          QnCtrl* __restrict__ ctrl, double* __restrict__ rel_trace, double* __restrict__ abs_trace, int step, double eps, double protect,
          int threshold) {
    if (ctrl->done) return;
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int row = gwarp; row < nhist * 3 + 2; row += nwarps) {       // the last two rows are ⟨δx,δg⟩ and ⟨δx,g⟩
        const float* p = partial + (int64_t)row * num_chunks;
        double s = 0.0;
        for (int i = lane; i < num_chunks; i += 32) s += (double)p[i];
        s = warp_sum_d(s);
        if (lane == 0) {
            dbuf[row] = s;
            if (row < nhist * 3) coef[(row % 3) * cap + row / 3] = (float)s;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        double n1 = 0.0, n2 = 0.0;
        for (int i = lane; i < norm_blocks; i += 32) {
            n1 += (double)norm_part[i];
            n2 += (double)norm_part[norm_blocks + i];
        }
        n1 = warp_sum_d(n1);
        n2 = warp_sum_d(n2);
        qn_decide(lane, n1, n2, ctrl, rel_trace, abs_trace, step, eps, protect, threshold);
    }
}

// ---- mesh-partitioned variant: local sums → (NCCL all-reduce of the fp64 buffer) → coefficients + stop rules -------------
// dbuf[row] = Σ_chunks partial[row], row = k*3+q, then ⟨δx,δg⟩, ⟨δx,g⟩ ; dbuf[3·nhist+2] = Σ‖g‖² partials ; dbuf[3·nhist+3] = Σ‖f‖² partials
__global__ void __launch_bounds__(256)
k_qn_fin1_local(int nhist, const float* __restrict__ partial, int num_chunks, double* __restrict__ dbuf, const float* __restrict__ norm_part,
                int norm_blocks, const QnCtrl* __restrict__ ctrl) {
    if (ctrl->done) return;
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int row = gwarp; row < nhist * 3 + 2; row += nwarps) {
        const float* p = partial + (int64_t)row * num_chunks;
        double s = 0.0;
        for (int i = lane; i < num_chunks; i += 32) s += (double)p[i];
        s = warp_sum_d(s);
        if (lane == 0) dbuf[row] = s;
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        double n1 = 0.0, n2 = 0.0;
        for (int i = lane; i < norm_blocks; i += 32) {
            n1 += (double)norm_part[i];
            n2 += (double)norm_part[norm_blocks + i];
        }
        n1 = warp_sum_d(n1);
        n2 = warp_sum_d(n2);
        if (lane == 0) { dbuf[3 * nhist + 2] = n1; dbuf[3 * nhist + 3] = n2; }
    }
}

__global__ void __launch_bounds__(256)
k_qn_fin1_global(int nhist, const double* __restrict__ dbuf, float* __restrict__ coef, int cap, QnCtrl* __restrict__ ctrl,
                 double* __restrict__ rel_trace, double* __restrict__ abs_trace, int step, double eps, double protect, int threshold) {
    if (ctrl->done) return;
    for (int row = threadIdx.x; row < nhist * 3; row += blockDim.x) coef[(row % 3) * cap + row / 3] = (float)dbuf[row];
    if (threadIdx.x < 32) qn_decide(threadIdx.x, dbuf[3 * nhist + 2], dbuf[3 * nhist + 3], ctrl, rel_trace, abs_trace, step, eps, protect, threshold);
}

// ---- start of a solve: update_0 = g_0 ; x_1 = x_0 + update_0 ------------------------------------------------
__global__ void __launch_bounds__(QN_THREADS)
k_qn_first(float* __restrict__ dx_upd, const float* __restrict__ g, float* __restrict__ x, float* __restrict__ xtrace_next,
           int num_chunks) {
    for (int chunk = blockIdx.x; chunk < num_chunks; chunk += gridDim.x) {
        const int64_t e0 = (int64_t)chunk * QN_CHUNK + threadIdx.x * 4;
        const float4 vg = *reinterpret_cast<const float4*>(g + e0);
        const float4 xv = *reinterpret_cast<const float4*>(x + e0);
        // update = −matvec(∅, ∅, g) = −(−g)
        const float4 upd = make_float4(-(-vg.x), -(-vg.y), -(-vg.z), -(-vg.w));
        const float4 xn = make_float4(xv.x + upd.x, xv.y + upd.y, xv.z + upd.z, xv.w + upd.w);
        *reinterpret_cast<float4*>(dx_upd + e0) = make_float4(xn.x - xv.x, xn.y - xv.y, xn.z - xv.z, xn.w - xv.w);
        *reinterpret_cast<float4*>(x + e0) = xn;
        if (xtrace_next != nullptr) *reinterpret_cast<float4*>(xtrace_next + e0) = xn;
    }
}

// g = fx − x ; δg = g − g_old ; norm partials — the generic-operator twin of solver_epilogue (layer.cuh)
__global__ void __launch_bounds__(QN_THREADS)
k_qn_post(int64_t numel, const float* __restrict__ fx, const float* __restrict__ x, float* __restrict__ g, float* __restrict__ dg,
          float* __restrict__ norm_part, const int* __restrict__ done) {
    __shared__ float smem[2 * (QN_THREADS / 32)];
    if (*done) return;
    float acc[2] = {0.f, 0.f};
    const int64_t i0 = ((int64_t)blockIdx.x * QN_THREADS + threadIdx.x) * 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t i = i0 + q;
        if (i < numel) {
            const float xv = x[i], gn = fx[i] - xv, fr = gn + xv;
            dg[i] = gn - g[i];
            g[i] = gn;
            acc[0] = fmaf(gn, gn, acc[0]);
            acc[1] = fmaf(fr, fr, acc[1]);
        }
    }
    block_sum<2, QN_THREADS / 32>(acc, smem);
    if (threadIdx.x == 0) {
        norm_part[blockIdx.x] = acc[0];
        norm_part[gridDim.x + blockIdx.x] = acc[1];
    }
}

__global__ void k_qn_ctrl_init(QnCtrl* ctrl, int* host_done) {
    ctrl->host_done = host_done;
    ctrl->done = 0; ctrl->improved = 0; ctrl->nstep = 0; ctrl->prot_break = 0; ctrl->stop_reason = 0;
    ctrl->best_step_rel = 0; ctrl->best_step_abs = 0; ctrl->pad_ = 0;
    ctrl->best_rel = 1e8; ctrl->best_abs = 1e8; ctrl->first_rel = 0.0; ctrl->s = 0.0; ctrl->p = 0.0;
}
