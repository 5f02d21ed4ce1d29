/* psignn_b200.h — C ABI of the B200-native PSI-GNN implicit message-passing solve.
 *
 * The reference (mnastorg/PSI-GNN) is pure Python and has no FFI of its own; its hot path is
 * entered through three Python call sites.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference root):
 *
 *   dirichlet/psignn/model.py:279-300   Function.forward          -> psi_layer_forward
 *   mixed/psignn/model.py:216-245       Function.forward (mixed)  -> psi_layer_forward (kind 1)
 *   dirichlet/dss/model.py:113-121      DSS layer                 -> psi_layer_forward (kind 2)
 *   dirichlet/dsgps/model.py:143-163    DSGPS recurrent step      -> psi_layer_forward (kind 3)
 *   mixed/dsgps/model.py:76-97          DSGPS step, mixed BCs     -> psi_layer_forward (kind 4)
 *   dirichlet/psignn/model.py:214       autograd.grad(f(H*),H*,y) -> psi_vjp_prepare / psi_vjp_apply
 *   dirichlet/psignn/model.py:157-167   residual_loss             -> psi_residual
 *   dirichlet/psignn/model.py:370-389   Encoder / Decoder         -> psi_encode / psi_decode
 *   utilities/solver.py:116-207         broyden                   -> psi_solver_broyden (+ step API)
 *   utilities/solver.py:215-293         anderson                  -> psi_solver_anderson
 *   utilities/solver.py:301-341         forward_iteration         -> psi_solver_picard
 *   model.py:342,360 remove_self_loops; :281 where(tags==1); :159-163 SparseTensor(...)
 *                                                                 -> psi_graph_create (done once)
 *
 * Conventions
 *   - every pointer argument marked "dev" is a BORROWED device pointer (e.g. into torch-owned
 *     storage) on the current CUDA device; the library owns only the handles it creates.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - all functions return 0 on success, non-zero on failure; psi_last_error() returns the
 *     message of the last failure on the calling thread.
 *   - a handle is not thread-safe; use one handle per (device, stream).  The layer weights live
 *     in one __constant__ block per process: one solve at a time per device.
 *   - all floating point is fp32 (the reference's dtype); latent width d is fixed at 10
 *     (every shipped config and checkpoint of the reference).
 */
#ifndef PSIGNN_B200_H
#define PSIGNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSI_LATENT_DIM 10

/* model kinds (layer variants sharing the fused layer kernel) */
#define PSI_KIND_DIRICHLET 0   /* dirichlet/psignn: edge attr 3, prb 2, gate, LayerNorm, Dirichlet clamp   */
#define PSI_KIND_MIXED     1   /* mixed/psignn:     + Neumann branch, prb 3, unit normals 2                 */
#define PSI_KIND_DSS       2   /* dirichlet/dss:    edge attr 1, b' 3, H += alpha*Psi, no LN, no clamp      */
#define PSI_KIND_DSGPS     3   /* dirichlet/dsgps:  GRU-style gate, Dirichlet clamp                         */
#define PSI_KIND_DSGPS_MIXED 4 /* mixed/dsgps:      + Neumann overwrite (phi_neumann, update_neumann), prb 3, unit normals 2 */

typedef struct psi_graph  psi_graph_t;    /* re-laid-out batch of meshes (destination-sorted, warp-sliced CSR) */
typedef struct psi_solver psi_solver_t;   /* fixed-point solver workspace (iterate, residual, U/V history)      */
typedef struct psi_comm   psi_comm_t;     /* NCCL communicator of a mesh-partitioned solve                      */

/* result block of a solve (mirrors the dict returned by the reference solvers) */
typedef struct psi_solve_stats {
    double  lowest;          /* best relative residual seen                       ("lowest")     */
    int32_t nstep;           /* step index of the best iterate                    ("nstep")      */
    int32_t steps_run;       /* number of solver steps actually executed                          */
    int32_t prot_break;      /* protective break taken                            ("prot_break") */
    int32_t stop_reason;     /* 0 threshold, 1 rel<eps, 2 plateau, 3 protective break             */
    int32_t f_evals;         /* operator evaluations                                               */
    int32_t launches;        /* CUDA kernels launched by this call                                 */
} psi_solve_stats_t;

int         psi_version(void);
const char* psi_last_error(void);
/* number of floats of the packed weight block (layout: psi_gnn_b200/weights.py mirrors csrc/weights.cuh) */
int         psi_weights_floats(void);
/* copy a packed weight block (device pointer) into the layer-constant bank, stream ordered */
int         psi_weights_upload(const float* dev_blob, int n_floats, void* stream);

/* ---- graph re-layout (once per batch) -------------------------------------------------------- */
int psi_graph_create(psi_graph_t** out, int64_t num_nodes, int64_t nnz,
                     const int64_t* dev_edge_index,   /* [2, nnz] (row, col) of A incl. diagonal      */
                     const float*   dev_edge_attr,    /* [nnz, attr_dim]                               */
                     int            attr_dim,         /* 3 (psignn, dsgps) or 1 (dss)                  */
                     const float*   dev_a_ij,         /* [nnz] stiffness coefficients, may be NULL     */
                     const float*   dev_tags,         /* [N, tag_dim] 0/1 floats, may be NULL          */
                     int            tag_dim,          /* 1 (dirichlet) or 3 (mixed one-hot)            */
                     const float*   dev_prb,          /* [N, prb_dim] second member                    */
                     int            prb_dim,          /* 2 or 3                                        */
                     const float*   dev_normals,      /* [N, 2] or NULL                                */
                     void* stream);
int psi_graph_destroy(psi_graph_t* g);
/* info[0]=N, [1]=E (off-diagonal edges), [2]=nnz, [3]=#dirichlet, [4]=#neumann,
 * [5]=slots 'to' list, [6]=slots 'from' list, [7]=bytes owned by the handle */
int psi_graph_info(const psi_graph_t* g, int64_t info[8]);

/* ---- mesh-partitioned solve (one large mesh split by node ranges over the ranks; replaces nothing in the reference, whose only
 * parallelism is PyG DataParallel at dirichlet/psignn/main.py:106 — this is BASELINE config 5) ------------------------------------
 * Local numbering of a rank: owned nodes [0, n_owned), then ghost nodes grouped by owning peer in the order of peer_ranks.
 * The graph is created over owned + ghost nodes; after psi_graph_set_partition the operator produces the owned rows only, ghost rows
 * of the iterate are refreshed from their owners before every operator evaluation (ncclSend/ncclRecv on the caller's stream), and the
 * Broyden inner products and norms are all-reduced (one exchange of 3(n-1)+4 fp64 sums per step).  The backward (VJP) solve exchanges
 * the ghost rows of S̄ between its two phases instead.  NCCL is resolved with dlopen at run time. */
int psi_comm_unique_id(char out[128]);                                   /* rank 0; broadcast the bytes to the other ranks */
int psi_comm_create(psi_comm_t** out, int rank, int world, const char id[128]);
int psi_comm_destroy(psi_comm_t* c);
int psi_graph_set_partition(psi_graph_t* g, psi_comm_t* comm, int64_t n_owned, int n_peers, const int32_t* peer_ranks /* host */,
                            const int64_t* send_counts /* host, rows per peer */, const int64_t* recv_counts /* host */,
                            const int32_t* dev_send_index /* device: owned local rows to send, concatenated per peer */, void* stream);
/* Peer-mapped mailboxes (NVLink/NVSwitch P2P through CUDA IPC): with them the halo rows are stored straight into the consumer's
 * memory by the producing rank and the Broyden inner products are exchanged and summed inside the reduction kernel — no NCCL call and
 * no packing kernel inside a solver step.  psi_part_mail_create allocates this rank's block and returns its 64-byte IPC handle and
 * ghost-row count; after an all-gather of both (host side, e.g. torch.distributed) psi_part_mail_open maps every rank's block.
 * remote_off[i] = position (rows) of this rank's rows in the ghost order of neighbour i.  Without these calls the partitioned solve
 * uses grouped ncclSend/ncclRecv + ncclAllReduce. */
int psi_part_mail_create(psi_graph_t* g, char handle_out[64], int64_t* total_recv_out);
int psi_part_mail_open(psi_graph_t* g, const char* handles /* [world][64] */, const int64_t* total_recvs /* [world] */,
                       const int64_t* remote_off /* [n_peers] */);
/* 1 if a device-side wait of this partition timed out (a peer did not arrive within 4 s): results of that solve are invalid */
int psi_part_error(const psi_graph_t* g);
/* refresh the ghost rows of a [N, width] fp32 array (width 2, 10 or 20) from their owners */
int psi_halo_exchange(psi_graph_t* g, float* dev_vec, int width, void* stream);

/* ---- one application of the layer and its transpose-Jacobian --------------------------------- */
int psi_layer_forward(const psi_graph_t* g, int kind, const float* dev_h, const float* dev_h0,
                      float* dev_out, void* stream);
/* n_layers unrolled applications in one call (DSS inference dirichlet/dss/model.py:113-123: one packed weight block per layer,
 * n_blobs = n_layers; DSGPS inference dirichlet/dsgps/model.py:143-160: the same block every step, n_blobs = 1).  dev_blobs holds n_blobs
 * blocks of psi_weights_floats() floats; the last uploaded block stays resident (its decoder is used by psi_decode). */
int psi_layers_unrolled(const psi_graph_t* g, int kind, const float* dev_blobs, int n_blobs, int n_layers, const float* dev_h,
                        const float* dev_h0, float* dev_work, float* dev_out, void* stream);
/* cache the linearisation point (ReLU masks, gate, LayerNorm statistics) for psi_vjp_apply */
int psi_vjp_prepare(psi_graph_t* g, int kind, const float* dev_hstar, const float* dev_h0, void* stream);
/* out = J^T y (+ grad if grad != NULL), J = d f / d h at the prepared point */
int psi_vjp_apply(psi_graph_t* g, int kind, const float* dev_y, const float* dev_grad,
                  float* dev_out, void* stream);

/* Parameter gradient of one application of f at the prepared point: theta_bar = (d f / d theta)^T y_bar, the walk that the reference's
 * loss.backward() makes from new_H_star = f(H*) to the parameters (dirichlet/psignn/model.py:204-205; mixed :146-147).  dev_out: flat
 * fp32 vector of psi_weights_floats() floats in the layout of the packed weight block (gradient of field X at the offset of X);
 * dev_jty (optional): J^T y_bar.  tab_dst/tab_y/tab_x [n_tab]: table of psi_gnn_b200/weights.py (parameter = sum over nodes of
 * record[tab_y] * record[tab_x], record layout from psi_pgrad_layout).  Deterministic (no atomics). */
int psi_param_grad(psi_graph_t* g, int kind, const float* dev_hstar, const float* dev_ybar, const int32_t* dev_tab_dst,
                   const int32_t* dev_tab_y, const int32_t* dev_tab_x, int n_tab, float* dev_out, float* dev_jty, void* stream);
/* Tangent of psi_param_grad along a direction dev_hdot of the frozen point: d/d eps theta_bar(H* + eps*hdot; y_bar) at eps = 0.  With
 * y_bar = v (Hutchinson probe) and hdot = J^T v it equals 1/2 grad_theta ||J^T v||^2, the double backward that the reference's
 * jac_loss_estimate(create_graph=True) + loss.backward() performs (dirichlet/psignn/model.py:207, :416-435; mixed :149, :355-374). */
int psi_param_grad_tangent(psi_graph_t* g, int kind, const float* dev_hstar, const float* dev_ybar, const float* dev_hdot,
                           const int32_t* dev_tab_dst, const int32_t* dev_tab_y, const int32_t* dev_tab_x, int n_tab, float* dev_out,
                           void* stream);
int psi_pgrad_layout(int32_t out[16]);
/* Backward of ONE unrolled layer of the DSS / DSGPS baselines at the layer's own input h — what autograd does for one step of the
 * reference's unrolled training forward (dirichlet/dss/model.py:83-91, dirichlet/dsgps/model.py:143-163, mixed/dsgps/model.py:76-97):
 * dev_hbar = (d f / d h)^T y_bar, dev_out = (d f / d theta)^T y_bar in packed-block layout (as psi_param_grad).  kind = PSI_KIND_DSS,
 * PSI_KIND_DSGPS or PSI_KIND_DSGPS_MIXED; the layer's weight block must be resident (psi_weights_upload); the table comes from
 * psi_gnn_b200/weights.py (baseline_grad_table; record layout = psi_pgrad_layout).  The gradient to h0 (the clamped Dirichlet rows
 * copy it) is y_bar on those rows and is left to the caller.  Deterministic (no atomics). */
int psi_layer_backward(psi_graph_t* g, int kind, const float* dev_h, const float* dev_ybar, const int32_t* dev_tab_dst,
                       const int32_t* dev_tab_y, const int32_t* dev_tab_x, int n_tab, float* dev_hbar, float* dev_out, void* stream);

/* ---- physics residual, encoder, decoder ------------------------------------------------------- */
/* r = A u - y (all nnz incl. diagonal); *dev_mean_sq = mean(r^2); dev_r may be NULL */
int psi_residual(const psi_graph_t* g, const float* dev_u, const float* dev_y, float* dev_r,
                 float* dev_mean_sq, void* stream);
/* out = A^T v   (used by the backward of the residual loss) */
int psi_spmv_t(const psi_graph_t* g, const float* dev_v, float* dev_out, void* stream);
/* flux form of the DSS residual (dirichlet/dss/model.py:137-145: F_bar = a_ij*(u_j-u_i) scatter_added by source row):
 * out_i = sum_{e=(i->j)} a_e (v_j - v_i); transpose != 0: its adjoint (the backward of the above).  Deterministic. */
int psi_flux(const psi_graph_t* g, const float* dev_v, float* dev_out, int transpose, void* stream);
int psi_encode(int64_t num_nodes, const float* dev_x, float* dev_h, void* stream);   /* MLP 1->d->d */
int psi_decode(int64_t num_nodes, const float* dev_h, float* dev_u, void* stream);   /* MLP d->d->1 */

/* ---- fixed-point solvers ----------------------------------------------------------------------- */
int psi_solver_create(psi_solver_t** out, int64_t numel /* N*d */, int max_threshold);
int psi_solver_destroy(psi_solver_t* s);
/* bytes of device memory currently owned by the solver workspace */
int64_t psi_solver_bytes(const psi_solver_t* s);
/* floats between consecutive vectors of the workspace (numel rounded up to 1024); the optional
 * xest_trace buffer passed to the Broyden entry points uses this stride */
int64_t psi_solver_stride(const psi_solver_t* s);
/* Optional per-kernel-class timing of the solver loops with CUDA events on the launching stream (used by bench.py for the
 * roofline).  Classes: 0 operator kernel(s) (fused layer, or the VJP pair), 1 k_qn_dots_tma, 2 k_qn_axpy_tma.
 * psi_solver_profile(s, 1) enables and resets the totals; psi_solver_profile_read fills, per class,
 * {launches, total ms, total algorithmic bytes}. */
int psi_solver_profile(psi_solver_t* s, int enable);
int psi_solver_profile_read(const psi_solver_t* s, double out[9]);

/* operator selector for the fused native loops */
#define PSI_OP_LAYER 0   /* x -> f(x; h0, graph)                 (forward solve,  model.py:189-193) */
#define PSI_OP_VJP   1   /* y -> J^T y + grad at prepared point  (backward solve, model.py:214-218) */

/* Broyden on g(x)=op(x)-x, whole batch as one vector, no line search (solver.py:116-207).
 * dev_aux: h0 for PSI_OP_LAYER, grad for PSI_OP_VJP.  dev_result receives the best iterate.
 * rel_trace/abs_trace: host arrays of threshold+1 doubles (padded with the lowest value, as the
 * reference does) or NULL.  dev_xtrace: optional [(threshold+1), stride] device buffer receiving
 * every iterate (xest_trace) or NULL; row stride psi_solver_stride(). */
int psi_solver_broyden(psi_solver_t* s, psi_graph_t* g, int kind, int op,
                       const float* dev_x0, const float* dev_aux, int threshold, double eps,
                       float* dev_result, psi_solve_stats_t* stats,
                       double* rel_trace, double* abs_trace, float* dev_xtrace, void* stream);

/* Step API of the same Broyden kernels for an arbitrary (e.g. Python) operator:
 *   begin(x0) ; fx = f(x) ; first(fx) ;  loop { fx = f(current x) ; step(fx) -> done? } ; finish */
int psi_broyden_begin(psi_solver_t* s, const float* dev_x0, int threshold, double eps, float* dev_xtrace, void* stream);
const float* psi_broyden_x(const psi_solver_t* s);                 /* device pointer of the current iterate */
int psi_broyden_first(psi_solver_t* s, const float* dev_fx, void* stream);
int psi_broyden_step(psi_solver_t* s, const float* dev_fx, int* done, void* stream);
int psi_broyden_finish(psi_solver_t* s, float* dev_result, psi_solve_stats_t* stats,
                       double* rel_trace, double* abs_trace, void* stream);
/* teacher-forced single rank-one update for parity tests: loads (x, gx, U[:n-1], V[:n-1]) and the new
 * (x_new, g_new), returns u_n, v_n, update_n computed by the production kernels */
int psi_broyden_forced_step(psi_solver_t* s, int n, const float* dev_x, const float* dev_gx,
                            const float* dev_xnew, const float* dev_gnew,
                            const float* dev_U /* [n-1, numel] */, const float* dev_V /* [n-1, numel] */,
                            float* dev_u, float* dev_v, float* dev_update, void* stream);

/* Anderson(m) (solver.py:215-293) and Picard iteration (solver.py:301-341) on the layer operator.
 * dev_xtrace: optional [(threshold+2), stride] device buffer receiving the reference's xest_trace (Anderson: x0, then the best
 * iterate so far after every step; Picard: every iterate) or NULL. */
int psi_solver_anderson(psi_solver_t* s, psi_graph_t* g, int kind, const float* dev_x0, const float* dev_h0,
                        int m, double lam, int threshold, double eps, double beta,
                        float* dev_result, psi_solve_stats_t* stats, double* rel_trace, double* abs_trace,
                        float* dev_xtrace, void* stream);
int psi_solver_picard(psi_solver_t* s, psi_graph_t* g, int kind, const float* dev_x0, const float* dev_h0,
                      int threshold, double eps, float* dev_result, psi_solve_stats_t* stats,
                      double* rel_trace, double* abs_trace, float* dev_xtrace, void* stream);
/* Step APIs of the same kernels for an arbitrary (e.g. Python) operator — the callable form of
 * anderson(f, x0, ...) / forward_iteration(f, z0, ...):
 *   begin(x0) ; loop { fx = f(psi_*_x()) ; psi_*_feed(fx, &done) } until done ; finish */
int psi_anderson_begin(psi_solver_t* s, const float* dev_x0, int m, double lam, int threshold, double eps, double beta,
                       float* dev_xtrace, void* stream);
const float* psi_anderson_x(const psi_solver_t* s);               /* device pointer of the point to evaluate next */
int psi_anderson_feed(psi_solver_t* s, const float* dev_fx, int* done, void* stream);
int psi_anderson_finish(psi_solver_t* s, float* dev_result, psi_solve_stats_t* stats, double* rel_trace, double* abs_trace,
                        void* stream);
/* teacher-forced single Anderson update for parity tests: window of n (<= m) vectors X, F ([n, numel] each) in; the production kernels
 * form the Gram matrix, solve the bordered system and mix: dev_xnew = beta*sum(alpha_i F_i) + (1-beta)*sum(alpha_i X_i), dev_alpha [n] */
int psi_anderson_forced_step(psi_solver_t* s, int m, int n, int slot, double lam, double beta, const float* dev_X, const float* dev_F,
                             float* dev_xnew, float* dev_alpha, void* stream);
int psi_picard_begin(psi_solver_t* s, const float* dev_z0, int threshold, double eps, float* dev_xtrace, void* stream);
const float* psi_picard_x(const psi_solver_t* s);
int psi_picard_feed(psi_solver_t* s, const float* dev_fx, int* done, void* stream);
int psi_picard_finish(psi_solver_t* s, float* dev_result, psi_solve_stats_t* stats, double* rel_trace, double* abs_trace,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PSIGNN_B200_H */
