// common.cuh — shared helpers for the PSI-GNN sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#define PSI_D 10                       // latent width (every shipped reference config)
#define PSI_NODE_BLOCK 128             // threads per CTA of the node-parallel kernels (4 warp-slices)
#define PSI_NUM_SMS_B200 148
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
