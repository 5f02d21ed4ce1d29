// anderson.cuh — Anderson acceleration and Picard iteration bookkeeping kernels.
//
// Reference: `anderson` (*/utilities/solver.py:215-293, stop_mode 'rel') and `forward_iteration`
// (solver.py:301-341).  The window X, F is stored as m contiguous vectors of `stride` floats; the
// (n+1)×(n+1) bordered system is solved on the device by one thread (n ≤ 8), so the loop has no host
// round trip; the stop flag lives in the same QnCtrl block the Broyden kernels use.
#pragma once
#include "common.cuh"
#include "broyden.cuh"

#define AND_MAX_M 8

// part[(i*n + j)*num_chunks + chunk] = Σ_chunk G_i·G_j,  G = F − X   (solver.py:250-251)
__global__ void __launch_bounds__(QN_THREADS)
k_and_gram(const float* __restrict__ X, const float* __restrict__ F, int64_t stride, int n, float* __restrict__ part, int num_chunks,
           const int* __restrict__ done) {
    __shared__ float smem[QN_THREADS / 32];
    if (*done) return;
    const int chunk = blockIdx.x;
    const int64_t e0 = (int64_t)chunk * QN_CHUNK + threadIdx.x * 4;
    float4 G[AND_MAX_M];
#pragma unroll
    for (int i = 0; i < AND_MAX_M; ++i) {
        if (i < n) {
            const float4 f = *reinterpret_cast<const float4*>(F + (int64_t)i * stride + e0);
            const float4 x = *reinterpret_cast<const float4*>(X + (int64_t)i * stride + e0);
            G[i] = make_float4(f.x - x.x, f.y - x.y, f.z - x.z, f.w - x.w);
        } else {
            G[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < AND_MAX_M; ++i) {
#pragma unroll
        for (int j = 0; j < AND_MAX_M; ++j) {
            if (i < n && j <= i) {
                float s = warp_sum(dot4(G[i], G[j], 0.f));
                if (lane == 0) smem[warp] = s;
                __syncthreads();
                if (threadIdx.x == 0) {
                    float t = 0.f;
#pragma unroll
                    for (int w = 0; w < QN_THREADS / 32; ++w) t += smem[w];
                    part[(int64_t)(i * n + j) * num_chunks + chunk] = t;
                    part[(int64_t)(j * n + i) * num_chunks + chunk] = t;
                }
                __syncthreads();
            }
        }
    }
}

// reduce the Gram partials, assemble H = [[0, 1ᵀ],[1, GGᵀ + λI]], solve H·a = e₀ (LU with partial pivoting, as
// torch.linalg.solve), alpha = a[1:n+1]   (solver.py:251-253)
__global__ void k_and_solve(const float* __restrict__ part, int num_chunks, int n, float lam, float* __restrict__ alpha,
                            const int* __restrict__ done) {
    __shared__ float H[(AND_MAX_M + 1) * (AND_MAX_M + 1)];
    if (*done) return;
    const int lane = threadIdx.x;
    const int dim = n + 1;
    for (int e = 0; e < n * n; ++e) {
        double s = 0.0;
        for (int c = lane; c < num_chunks; c += 32) s += (double)part[(int64_t)e * num_chunks + c];
        s = warp_sum_d(s);
        if (lane == 0) {
            const int i = e / n, j = e % n;
            H[(i + 1) * dim + (j + 1)] = (float)s + (i == j ? lam : 0.f);
        }
    }
    if (lane == 0) {
        H[0] = 0.f;
        for (int i = 1; i < dim; ++i) { H[i] = 1.f; H[i * dim] = 1.f; }
        float y[AND_MAX_M + 1];
        for (int i = 0; i < dim; ++i) y[i] = (i == 0) ? 1.f : 0.f;
        for (int c = 0; c < dim; ++c) {
            int piv = c;
            float best = fabsf(H[c * dim + c]);
            for (int r = c + 1; r < dim; ++r)
                if (fabsf(H[r * dim + c]) > best) { best = fabsf(H[r * dim + c]); piv = r; }
            if (piv != c) {
                for (int q = 0; q < dim; ++q) { const float t = H[c * dim + q]; H[c * dim + q] = H[piv * dim + q]; H[piv * dim + q] = t; }
                const float t = y[c]; y[c] = y[piv]; y[piv] = t;
            }
            const float d = H[c * dim + c];
            for (int r = c + 1; r < dim; ++r) {
                const float f = __fdiv_rn(H[r * dim + c], d);
                for (int q = c; q < dim; ++q) H[r * dim + q] = fmaf(-f, H[c * dim + q], H[r * dim + q]);
                y[r] = fmaf(-f, y[c], y[r]);
            }
        }
        for (int r = dim - 1; r >= 0; --r) {
            float t = y[r];
            for (int q = r + 1; q < dim; ++q) t = fmaf(-H[r * dim + q], y[q], t);
            y[r] = __fdiv_rn(t, H[r * dim + r]);
        }
        for (int i = 0; i < n; ++i) alpha[i] = y[i + 1];
    }
}

// X[slot] = β·Σ α_i F_i + (1−β)·Σ α_i X_i   (solver.py:255)
__global__ void __launch_bounds__(QN_THREADS)
k_and_mix(float* __restrict__ X, const float* __restrict__ F, int64_t stride, int n, int slot, const float* __restrict__ alpha, float beta,
          int num_chunks, const int* __restrict__ done) {
    if (*done) return;
    float a[AND_MAX_M];
#pragma unroll
    for (int i = 0; i < AND_MAX_M; ++i) a[i] = (i < n) ? alpha[i] : 0.f;
    const float omb = 1.0f - beta;
    for (int chunk = blockIdx.x; chunk < num_chunks; chunk += gridDim.x) {
        const int64_t e0 = (int64_t)chunk * QN_CHUNK + threadIdx.x * 4;
        float4 sf = make_float4(0.f, 0.f, 0.f, 0.f), sx = sf;
#pragma unroll
        for (int i = 0; i < AND_MAX_M; ++i) {
            if (i < n) {
                const float4 f = *reinterpret_cast<const float4*>(F + (int64_t)i * stride + e0);
                const float4 x = *reinterpret_cast<const float4*>(X + (int64_t)i * stride + e0);
                sf.x = fmaf(a[i], f.x, sf.x); sf.y = fmaf(a[i], f.y, sf.y); sf.z = fmaf(a[i], f.z, sf.z); sf.w = fmaf(a[i], f.w, sf.w);
                sx.x = fmaf(a[i], x.x, sx.x); sx.y = fmaf(a[i], x.y, sx.y); sx.z = fmaf(a[i], x.z, sx.z); sx.w = fmaf(a[i], x.w, sx.w);
            }
        }
        const float4 o = make_float4(beta * sf.x + omb * sx.x, beta * sf.y + omb * sx.y, beta * sf.z + omb * sx.z, beta * sf.w + omb * sx.w);
        *reinterpret_cast<float4*>(X + (int64_t)slot * stride + e0) = o;
    }
}

// gx = F[slot] − X[slot]: partial sums of ‖gx‖² and ‖F[slot]‖²   (solver.py:258-260)
__global__ void __launch_bounds__(QN_THREADS)
k_and_post(const float* __restrict__ xs, const float* __restrict__ fs, const float* __restrict__ unused, float* __restrict__ norm_part,
           int num_chunks, const int* __restrict__ done) {
    __shared__ float smem[2 * (QN_THREADS / 32)];
    (void)unused;
    if (*done) return;
    const int64_t e0 = (int64_t)blockIdx.x * QN_CHUNK + threadIdx.x * 4;
    const float4 f = *reinterpret_cast<const float4*>(fs + e0);
    const float4 x = *reinterpret_cast<const float4*>(xs + e0);
    const float4 gx = make_float4(f.x - x.x, f.y - x.y, f.z - x.z, f.w - x.w);
    float acc[2];
    acc[0] = dot4(gx, gx, 0.f);
    acc[1] = dot4(f, f, 0.f);
    block_sum<2, QN_THREADS / 32>(acc, smem);
    if (threadIdx.x == 0) {
        norm_part[blockIdx.x] = acc[0];
        norm_part[num_chunks + blockIdx.x] = acc[1];
    }
}

__global__ void k_and_fin(const float* __restrict__ norm_part, int norm_blocks, QnCtrl* __restrict__ ctrl, double* __restrict__ rel_trace,
                          double* __restrict__ abs_trace, int k, double eps) {
    if (ctrl->done) return;
    const int lane = threadIdx.x;
    double n1 = 0.0, n2 = 0.0;
    for (int i = lane; i < norm_blocks; i += 32) { n1 += (double)norm_part[i]; n2 += (double)norm_part[norm_blocks + i]; }
    n1 = warp_sum_d(n1);
    n2 = warp_sum_d(n2);
    if (lane == 0) {
        const double absd = (double)(float)sqrt(n1);
        const double rel = absd / (1e-5 + (double)(float)sqrt(n2));
        rel_trace[k - 2] = rel;
        abs_trace[k - 2] = absd;
        int improved = 0;
        if (rel < ctrl->best_rel) { ctrl->best_rel = rel; ctrl->best_step_rel = k; improved = 1; }
        if (absd < ctrl->best_abs) { ctrl->best_abs = absd; ctrl->best_step_abs = k; }
        ctrl->improved = improved;
        ctrl->nstep = k;
        if (rel < eps) { ctrl->done = 1; ctrl->stop_reason = 1; }
    }
}

// lowest_xest = X[slot].clone() when this step improved (solver.py:268-269).  `k` is the step this launch belongs to: launches
// queued behind the stop see the control block of the LAST executed step (k_and_fin no longer runs) and must not copy anything.
__global__ void __launch_bounds__(QN_THREADS)
k_and_keep(const float* __restrict__ xs, float* __restrict__ best, int num_chunks, QnCtrl* __restrict__ ctrl, int k) {
    if (!ctrl->improved || ctrl->nstep != k) return;
    for (int chunk = blockIdx.x; chunk < num_chunks; chunk += gridDim.x) {
        const int64_t e0 = (int64_t)chunk * QN_CHUNK + threadIdx.x * 4;
        *reinterpret_cast<float4*>(best + e0) = *reinterpret_cast<const float4*>(xs + e0);
    }
}

// Picard: evaluation e produced z_{e+1} = f(z_e) with partials of ‖z_{e+1} − z_e‖² and ‖z_{e+1}‖²; decide whether the
// reference's `while rel > eps and ite < threshold` continues (solver.py:314-331; norms are fp32 tensors there).
__global__ void k_picard_fin(const float* __restrict__ norm_part, int norm_blocks, QnCtrl* __restrict__ ctrl, double* __restrict__ rel_trace,
                             double* __restrict__ abs_trace, int e, float eps, int threshold) {
    if (ctrl->done) return;
    const int lane = threadIdx.x;
    double n1 = 0.0, n2 = 0.0;
    for (int i = lane; i < norm_blocks; i += 32) { n1 += (double)norm_part[i]; n2 += (double)norm_part[norm_blocks + i]; }
    n1 = warp_sum_d(n1);
    n2 = warp_sum_d(n2);
    if (lane == 0) {
        const float absf = (float)sqrt(n1);
        const float relf = __fdiv_rn(absf, (float)sqrt(n2));
        rel_trace[e] = (double)relf;
        abs_trace[e] = (double)absf;
        ctrl->best_rel = (double)relf;
        ctrl->nstep = e + 1;
        if (!(relf > eps) || e >= threshold) { ctrl->done = 1; ctrl->stop_reason = (e >= threshold && relf > eps) ? 0 : 1; }
    }
}
