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

struct LayerWeights {
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
    float pad_[2];
};

static_assert(sizeof(EdgeMLP) == 350 * 4, "EdgeMLP layout");
static_assert(sizeof(LayerWeights) % 8 == 0, "LayerWeights size");

__constant__ LayerWeights cW;   // single translation unit (psignn_b200.cu)

template <int WHICH> __device__ __forceinline__ const EdgeMLP& edge_mlp();
template <> __device__ __forceinline__ const EdgeMLP& edge_mlp<0>() { return cW.to; }
template <> __device__ __forceinline__ const EdgeMLP& edge_mlp<1>() { return cW.from; }
template <> __device__ __forceinline__ const EdgeMLP& edge_mlp<2>() { return cW.neu; }
