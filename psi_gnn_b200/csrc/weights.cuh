// weights.cuh — packed weight block of one message-passing layer, held in __constant__ memory.
//
// With fully unrolled loops every weight becomes a constant-bank operand of an FFMA
// (FFMA R, R, c[3][imm], R): one instruction per multiply-add, no load instructions, no shared
// memory traffic.  ncu justifies FMA over tensor cores for this path: the contractions are
// [E,23]x[23,10] and [N,32]x[32,10] in fp32 with a 1e-5 parity bar.
//
// The Python side (psi_gnn_b200/weights.py) packs nn.Module parameters into exactly this layout.
#pragma once
#include "common.cuh"

// Edge MLP  Linear(2d+A, d) -> ReLU -> Linear(d, d)   (reference MLP inside Phi_to / Phi_from,
// dirichlet/psignn/model.py:316-368).  The first layer is stored split by input block.
struct EdgeMLP {
    float W1i[PSI_D][PSI_D];   // columns 0..d-1   (multiplies h_i, the aggregation target)
    float W1j[PSI_D][PSI_D];   // columns d..2d-1  (multiplies h_j, the neighbour)
    float W1a[PSI_D][3];       // columns 2d..     (edge attributes; DSS uses column 0 only)
    float b1[PSI_D];
    float W2[PSI_D][PSI_D];
    float b2[PSI_D];
};  // 350 floats

struct __align__(16) LayerWeights {
    EdgeMLP to, from, neu;                                   // phi_to, phi_from, phi_neumann
    float gate_w[33];                                        // alpha: Linear(3d+s, 1)   (model.py:275)
    float gate_b;
    float up_W1[PSI_D][33];                                  // update / Psi: Linear(3d+s, d)
    float up_b1[PSI_D];
    float up_W2[PSI_D][PSI_D];
    float up_b2[PSI_D];
    float un_W1[PSI_D][25];                                  // update_neumann: Linear(2d+3+2, d)
    float un_b1[PSI_D];
    float un_W2[PSI_D][PSI_D];
    float un_b2[PSI_D];
    float ln_g[PSI_D];                                       // LayerNorm affine
    float ln_b[PSI_D];
    float gz_W[PSI_D][33]; float gz_b[PSI_D];                // DSGPS z_k         Linear(3d+s, d), s = 2 (dirichlet) or 3 (mixed)
    float gr_W[PSI_D][33]; float gr_b[PSI_D];                // DSGPS r_k
    float gc_W[PSI_D][33]; float gc_b[PSI_D];                // DSGPS correction
    float enc_W1[PSI_D]; float enc_b1[PSI_D];                // encoder 1 -> d -> d
    float enc_W2[PSI_D][PSI_D]; float enc_b2[PSI_D];
    float dec_W1[PSI_D][PSI_D]; float dec_b1[PSI_D];         // decoder d -> d -> 1
    float dec_W2[PSI_D]; float dec_b2;
    float dss_alpha;                                         // DSS constant step (config["alpha"])
    float pad_[4];                                           // sizeof = 3200 floats: a multiple of 16 bytes
};

static_assert(sizeof(EdgeMLP) == 350 * 4, "EdgeMLP layout");
static_assert(sizeof(LayerWeights) == 3200 * 4, "LayerWeights size");

// Transposed ([input][output]) copies of the matrices the forward kernels contract over their INPUT index: two adjacent output
// channels (o, o+1) then form one 8-byte constant operand of a packed FFMA2.  Packed by psi_gnn_b200/weights.py behind the
// LayerWeights block (same blob); the biases are vectors and are read pairwise from LayerWeights itself.
struct EdgeMLPT {
    float W1iT[PSI_D][PSI_D];
    float W1jT[PSI_D][PSI_D];
    float W1aT[3][PSI_D];
    float W2T[PSI_D][PSI_D];
};  // 330 floats
struct __align__(16) LayerWeightsT {
    EdgeMLPT to, from, neu;
    float up_W1T[33][PSI_D];
    float up_W2T[PSI_D][PSI_D];
    float un_W1T[25][PSI_D];
    float un_W2T[PSI_D][PSI_D];
    float gzT[33][PSI_D];
    float grT[33][PSI_D];
    float gcT[33][PSI_D];
};
static_assert(sizeof(EdgeMLPT) == 330 * 4, "EdgeMLPT layout");
static_assert(sizeof(LayerWeightsT) == 2760 * 4, "LayerWeightsT layout");

// one contiguous constant block (a weight upload — per layer for the unrolled DSS baseline — is a single 23.8 KB copy)
struct __align__(16) LayerBlock {
    LayerWeights w;
    LayerWeightsT t;
};
__constant__ LayerBlock cB;     // single translation unit (psignn_b200.cu)
#define cW (cB.w)
#define cWT (cB.t)

template <int WHICH> __device__ __forceinline__ const EdgeMLPT& edge_mlp_t();
template <> __device__ __forceinline__ const EdgeMLPT& edge_mlp_t<0>() { return cWT.to; }
template <> __device__ __forceinline__ const EdgeMLPT& edge_mlp_t<1>() { return cWT.from; }
template <> __device__ __forceinline__ const EdgeMLPT& edge_mlp_t<2>() { return cWT.neu; }

template <int WHICH> __device__ __forceinline__ const EdgeMLP& edge_mlp();
template <> __device__ __forceinline__ const EdgeMLP& edge_mlp<0>() { return cW.to; }
template <> __device__ __forceinline__ const EdgeMLP& edge_mlp<1>() { return cW.from; }
template <> __device__ __forceinline__ const EdgeMLP& edge_mlp<2>() { return cW.neu; }
