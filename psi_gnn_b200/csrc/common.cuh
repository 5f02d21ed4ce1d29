// common.cuh — shared helpers for the PSI-GNN sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#define PSI_D 10                       // latent width (every shipped reference config)
#define PSI_NODE_BLOCK 128             // threads per CTA of the node-parallel kernels (4 warp-slices)
#define PSI_NUM_SMS_B200 148
#define PSI_QPITCH 12                   // floats per row of the gathered per-node arrays (10 + 2 pad: 16-byte aligned rows)
#define PSI_STAGE_FLOATS (32 * PSI_QPITCH) // per-warp shared-memory stage of the cooperative row gather
#ifndef PSI_OP_MIN_CTAS
#define PSI_OP_MIN_CTAS 5                // resident CTAs per SM the operator kernels are compiled for (register cap 65536 / (128·5) = 102 → 96)
#endif

extern thread_local std::string g_psi_err;

#define PSI_FAIL(msg)                                                                   \
    do {                                                                                \
        g_psi_err = std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + msg; \
        return -1;                                                                      \
    } while (0)

#define PSI_CK(call)                                                                    \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            g_psi_err = std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + \
                        #call + " -> " + cudaGetErrorString(e__);                       \
            return -1;                                                                  \
        }                                                                               \
    } while (0)

#define PSI_CK_LAUNCH() PSI_CK(cudaGetLastError())

// Stream-ordered allocation from the device's default memory pool (kept cached: release threshold = max), so that the
// per-batch graph re-layout does not pay cudaMalloc/cudaFree device synchronisations.
static inline cudaError_t psi_malloc_async(void** p, size_t bytes, cudaStream_t st) {
    static thread_local int pool_ready_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (pool_ready_dev != dev) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t thr = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        pool_ready_dev = dev;
    }
    return cudaMallocAsync(p, bytes > 0 ? bytes : 16, st);
}
static inline void psi_free_async(void* p, cudaStream_t st) {
    if (p) cudaFreeAsync(p, st);
}

static inline int64_t round_up64(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// ---- row access: a latent row is 10 contiguous floats (40 B, 8-byte aligned) -----------------
__device__ __forceinline__ void load_row(const float* __restrict__ base, int64_t node, float (&v)[PSI_D]) {
    const float2* p = reinterpret_cast<const float2*>(base + node * PSI_D);
#pragma unroll
    for (int q = 0; q < PSI_D / 2; ++q) {
        float2 t = __ldg(p + q);
        v[2 * q] = t.x;
        v[2 * q + 1] = t.y;
    }
}
// same, but through the coherent path (for buffers written earlier in the same kernel chain is fine
// either way; this one is used where the compiler must not assume read-only)
__device__ __forceinline__ void load_row_rw(const float* base, int64_t node, float (&v)[PSI_D]) {
    const float2* p = reinterpret_cast<const float2*>(base + node * PSI_D);
#pragma unroll
    for (int q = 0; q < PSI_D / 2; ++q) {
        float2 t = p[q];
        v[2 * q] = t.x;
        v[2 * q + 1] = t.y;
    }
}
__device__ __forceinline__ void store_row(float* base, int64_t node, const float (&v)[PSI_D]) {
    float2* p = reinterpret_cast<float2*>(base + node * PSI_D);
#pragma unroll
    for (int q = 0; q < PSI_D / 2; ++q) p[q] = make_float2(v[2 * q], v[2 * q + 1]);
}

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2 — two IEEE fp32 operations per issue slot) --------------------
// The operator kernels are issue-bound (ncu: profiles/r02_a_operator.md), so the MLP contractions pair two output channels per
// instruction.  Each half is an ordinary round-to-nearest fp32 fma/add/mul: results are bit-identical to the scalar chains.
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(f2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2 ffma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 fadd2(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 fmul2(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ void unpack10(const f2 (&v)[PSI_D / 2], float (&x)[PSI_D]) {
#pragma unroll
    for (int q = 0; q < PSI_D / 2; ++q) upk(v[q], x[2 * q], x[2 * q + 1]);
}
__device__ __forceinline__ void pack10(const float (&x)[PSI_D], f2 (&v)[PSI_D / 2]) {
#pragma unroll
    for (int q = 0; q < PSI_D / 2; ++q) v[q] = pk(x[2 * q], x[2 * q + 1]);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// 16-byte asynchronous copy global → shared (LDGSTS): no register staging, completion tracked per thread in commit groups
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// a 10-float row stored at the 12-float pitch of the gathered arrays (pad written as zeros)
__device__ __forceinline__ void store_row12(float* base, int64_t node, const float (&v)[PSI_D]) {
    float4* p = reinterpret_cast<float4*>(base + node * PSI_QPITCH);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
    p[2] = make_float4(v[8], v[9], 0.f, 0.f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block-wide sum of NV values per thread; result valid in thread 0.
template <int NV, int NWARPS>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* smem /* NV*NWARPS floats */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float s = warp_sum(v[i]);
        if (lane == 0) smem[i * NWARPS + warp] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < NWARPS; ++w) s += smem[i * NWARPS + w];
            v[i] = s;
        }
    }
}
