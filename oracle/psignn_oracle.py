"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the PSI-GNN hot path.

Plain ``torch`` (CPU, fp32 or fp64) restatement of the reference algorithm for
the implicit message-passing solve, written against ``state_dict`` key names so
that the same weights drive the reference, this oracle and the CUDA path.
Every function cites the reference lines it follows.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the *unmodified*
reference (imported from /root/reference behind ``oracle/ref_shim.py``) on
seeded synthetic meshes with the shipped checkpoints and with random-init
weights and stores inputs + outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against those vectors (and,
where /root/reference is present, against the live reference).  The third-party
semantics the reference relies on (PyG ``MessagePassing`` add-aggregation,
``remove_self_loops``, ``torch_sparse`` SpMV) are pinned nowhere by the
reference itself; they are restated here as gather / index_add.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
(``psi_gnn_b200``) never does.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# ---------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------

def _lin(P: Params, key: str, x: Tensor) -> Tensor:
    return F.linear(x, P[key + ".weight"], P[key + ".bias"])


def mlp2(P: Params, prefix: str, x: Tensor) -> Tensor:
    """Linear → ReLU → Linear (reference ``MLP``, dirichlet/psignn/model.py:316-332)."""
    return _lin(P, prefix + ".2", torch.relu(_lin(P, prefix + ".0", x)))


def offdiag(edge_index: Tensor, *edge_tensors: Tensor):
    """``remove_self_loops`` (call sites model.py:342,360)."""
    keep = edge_index[0] != edge_index[1]
    return (edge_index[:, keep],) + tuple(t[keep] for t in edge_tensors)


def phi(P: Params, prefix: str, h: Tensor, ei: Tensor, attr: Tensor, to: bool) -> Tensor:
    """Σ_e MLP(cat[h_i, h_j, a_e]) (reference ``Phi_to``/``Phi_from``, model.py:334-368).

    ``to=True``  : flow source_to_target — i = ei[1], j = ei[0], summed at ei[1].
    ``to=False`` : flow target_to_source — i = ei[0], j = ei[1], summed at ei[0].
    ``ei``/``attr`` must already be free of self loops.
    """
    i, j = (ei[1], ei[0]) if to else (ei[0], ei[1])
    msg = mlp2(P, prefix + ".mlp.mlp", torch.cat([h[i], h[j], attr], 1))
    return torch.zeros(h.size(0), msg.size(1), dtype=h.dtype).index_add(0, i, msg)


# ---------------------------------------------------------------------------
# f_theta : one message-passing layer
# ---------------------------------------------------------------------------

def f_dirichlet(P: Params, h: Tensor, h0: Tensor, batch, prefix: str = "deqdss.f", n_layers: int = 1) -> Tensor:
    """``Function.forward`` (dirichlet/psignn/model.py:279-300)."""
    ei, attr = offdiag(batch.edge_index, batch.edge_attr)
    dirichlet = torch.where(batch.tags.reshape(-1) == 1)[0]
    for k in range(n_layers):
        to = phi(P, f"{prefix}.phi_to_list.{k}", h, ei, attr, True)
        fr = phi(P, f"{prefix}.phi_from_list.{k}", h, ei, attr, False)
        c = torch.cat([h, to, fr, batch.prb_data], 1)
        alpha = torch.sigmoid(_lin(P, f"{prefix}.alpha.0", c))
        upd = alpha * mlp2(P, f"{prefix}.update_list.{k}.mlp", c)
        h = h + upd
        if k == n_layers - 1:
            h = F.layer_norm(h, (h.size(1),), P[f"{prefix}.laynorm.weight"], P[f"{prefix}.laynorm.bias"], 1e-5)
        h = h.index_copy(0, dirichlet, h0[dirichlet])
    return h


def f_mixed(P: Params, h: Tensor, h0: Tensor, batch, prefix: str = "deqdss.f", n_layers: int = 1) -> Tensor:
    """Mixed ``Function.forward`` (mixed/psignn/model.py:216-245).

    Keeps the reference quirk that ``h`` is never reassigned inside the layer loop
    (only the last layer's output survives); shipped configs use ``n_layers=1``.
    """
    ei, attr = offdiag(batch.edge_index, batch.edge_attr)
    dirichlet = torch.where(batch.tags[:, 1] == 1)[0]
    neumann = torch.where(batch.tags[:, 2] == 1)[0]
    out = h
    for k in range(n_layers):
        to = phi(P, f"{prefix}.phi_to_list.{k}", h, ei, attr, True)
        fr = phi(P, f"{prefix}.phi_from_list.{k}", h, ei, attr, False)
        neu = phi(P, f"{prefix}.phi_neumann", h, ei, attr, False)
        c = torch.cat([h, to, fr, batch.prb_data], 1)
        alpha = torch.sigmoid(_lin(P, f"{prefix}.alpha.0", c))
        upd = alpha * mlp2(P, f"{prefix}.update_list.{k}.mlp", c)
        cn = torch.cat([h, neu, batch.prb_data, batch.unit_normal_vector], 1)
        upd_n = mlp2(P, f"{prefix}.update_neumann.mlp", cn)
        out = h + upd
        out = out.index_copy(0, neumann, upd_n[neumann])
        if k == n_layers - 1:
            out = F.layer_norm(out, (h.size(1),), P[f"{prefix}.laynorm.weight"], P[f"{prefix}.laynorm.bias"], 1e-5)
        out = out.index_copy(0, dirichlet, h0[dirichlet])
    return out


def dss_layer(P: Params, k: int, H: Tensor, batch, alpha: float) -> Tensor:
    """One DSS update ``H += alpha * Psi_k(cat[H, to, from, b'])`` (dirichlet/dss/model.py:113-121)."""
    ei, attr = offdiag(batch.edge_index, batch.a_ij_norm)
    to = phi(P, f"phi_to_list.{k}", H, ei, attr, True)
    fr = phi(P, f"phi_from_list.{k}", H, ei, attr, False)
    c = torch.cat([H, to, fr, batch.b_prime_norm], 1)
    return H + alpha * mlp2(P, f"psi_list.{k}.mlp.mlp", c)


def dss_inference(P: Params, batch, k_layers: int, alpha: float, latent_dim: int = 10) -> Tensor:
    """``DeepStatisticalSolver.inference`` (dirichlet/dss/model.py:106-127)."""
    H = torch.zeros(batch.num_nodes, latent_dim, dtype=batch.x.dtype)
    for k in range(k_layers):
        H = dss_layer(P, k, H, batch, alpha)
    return mlp2(P, f"decoder_list.{k_layers - 1}.mlp.mlp", H)


def dsgps_layer(P: Params, H: Tensor, H0: Tensor, batch) -> Tensor:
    """One DSGPS recurrent step (dirichlet/dsgps/model.py:143-163)."""
    ei, attr = offdiag(batch.edge_index, batch.edge_attr)
    dirichlet = torch.where(batch.tags.reshape(-1) == 1)[0]
    to = phi(P, "phi_to", H, ei, attr, True)
    fr = phi(P, "phi_from", H, ei, attr, False)
    c = torch.cat([H, to, fr, batch.prb_data], 1)
    z = torch.sigmoid(_lin(P, "z_k.mlp.0", c))
    r = torch.sigmoid(_lin(P, "r_k.mlp.0", c))
    corr = torch.tanh(_lin(P, "correction.mlp.0", torch.cat([r * H, to, fr, batch.prb_data], 1)))
    Hn = H + z * corr
    return Hn.index_copy(0, dirichlet, H0[dirichlet])


def dsgps_inference(P: Params, batch, k_steps: int) -> Tensor:
    """``ModelDSGPS.inference`` (dirichlet/dsgps/model.py:133-163)."""
    H0 = encoder(P, batch.x)
    H = H0
    for _ in range(k_steps):
        H = dsgps_layer(P, H, H0, batch)
    return decoder(P, H)


def dsgps_layer_mixed(P: Params, H: Tensor, H0: Tensor, batch) -> Tensor:
    """One mixed-boundary DSGPS step (mixed/dsgps/model.py:76-97): interior GRU-style update, Neumann rows overwritten by
    ``update_neumann(cat[H, ΣΦ_neumann, prb, n̂])``, Dirichlet rows (tags[:,1]) copied from ``H0`` — in that order."""
    ei, attr = offdiag(batch.edge_index, batch.edge_attr)
    to = phi(P, "phi_to", H, ei, attr, True)
    fr = phi(P, "phi_from", H, ei, attr, False)
    ne = phi(P, "phi_neumann", H, ei, attr, False)
    c = torch.cat([H, to, fr, batch.prb_data], 1)
    z = torch.sigmoid(_lin(P, "z_k.mlp.0", c))
    r = torch.sigmoid(_lin(P, "r_k.mlp.0", c))
    corr = torch.tanh(_lin(P, "correction.mlp.0", torch.cat([r * H, to, fr, batch.prb_data], 1)))
    upd = mlp2(P, "update_neumann.mlp", torch.cat([H, ne, batch.prb_data, batch.unit_normal_vector], 1))
    Hn = torch.where((batch.tags[:, 2] == 1)[:, None], upd, H + z * corr)
    return torch.where((batch.tags[:, 1] == 1)[:, None], H0, Hn)


def dss_flux_residual(U: Tensor, batch) -> Tensor:
    """flux-form residual of DSS (dirichlet/dss/model.py:129-148): mean((p1 + Σ_{e=(i→j)} a_e (u_j − u_i))²)"""
    frm, to = batch.edge_index[0], batch.edge_index[1]
    y = batch.b_prime
    p1 = (1 - y[:, 1:2]) * (-y[:, 0:1]) + y[:, 1:2] * (U - y[:, 2:3])
    flux = torch.zeros_like(U).index_add(0, frm, batch.a_ij.reshape(-1, 1) * (U[to] - U[frm]))
    return torch.mean((p1 + flux) ** 2)


def dss_training_forward(P: Params, batch, k_layers: int, alpha: float, gamma: float, latent_dim: int = 10):
    """``DeepStatisticalSolver.forward`` (dirichlet/dss/model.py:59-104): k unrolled layers from H = 0, per-layer decoders, the flux
    residual of every state weighted by gamma^(k−1−layer).  Returns (train_loss, U_last, residual of the last state)."""
    H = torch.zeros(batch.num_nodes, latent_dim, dtype=batch.a_ij.dtype)
    total = None
    for k in range(k_layers):
        H = dss_layer(P, k, H, batch, alpha)
        U = mlp2(P, f"decoder_list.{k}.mlp.mlp", H)
        res = dss_flux_residual(U, batch)
        term = res * gamma ** (k_layers - k - 1)
        total = term if total is None else total + term
    return total, U, res


def dsgps_training_forward(P: Params, batch, k_steps: int, gamma: float, mixed: bool = False):
    """``ModelDSGPS.forward`` (dirichlet/dsgps/model.py:48-131, mixed/dsgps/model.py:48-131): k recurrent steps from H0 = encoder(x);
    per step the residual of the decoded state weighted by gamma^(k−1−step) plus an encoder and an autoencoder loss.  The dirichlet
    reference freezes the decoder (resp. encoder) PARAMETERS for those two losses (the states keep their graph); the mixed reference
    detaches the STATES instead.  Returns (train_loss, U_last, residual of the last state)."""
    enc_keys = [k for k in P if k.startswith("autoencoder.encoder.")]
    dec_keys = [k for k in P if k.startswith("autoencoder.decoder.")]
    P_dec_frozen = {**P, **{k: P[k].detach() for k in dec_keys}}
    P_enc_frozen = {**P, **{k: P[k].detach() for k in enc_keys}}
    H0 = encoder(P, batch.x)
    H = H0
    total = None
    for step in range(k_steps):
        H = dsgps_layer_mixed(P, H, H0, batch) if mixed else dsgps_layer(P, H, H0, batch)
        U = decoder(P, H)
        res = residual_loss(U, batch)
        if mixed:
            u_d, h_d = U.detach(), H.detach()
            l_enc = F.mse_loss(encoder(P, u_d), h_d)
            l_auto = F.mse_loss(decoder(P, encoder(P, u_d).detach()), u_d)
        else:
            l_enc = F.mse_loss(encoder(P_dec_frozen, decoder(P_dec_frozen, H)), H)          # autoencoder(H, sens="latent")
            l_auto = F.mse_loss(decoder(P_enc_frozen, encoder(P_enc_frozen, U)), U)         # autoencoder(U, sens="physics")
        term = res * gamma ** (k_steps - step - 1) + l_enc + l_auto
        total = term if total is None else total + term
    return total, U, res


def encoder(P: Params, x: Tensor) -> Tensor:
    """MLP 1→d→d (model.py:370-378)."""
    return mlp2(P, "autoencoder.encoder.mlp.mlp", x)


def decoder(P: Params, h: Tensor) -> Tensor:
    """MLP d→d→1 (model.py:380-389)."""
    return mlp2(P, "autoencoder.decoder.mlp.mlp", h)


def residual_loss(u: Tensor, batch) -> Tensor:
    """mean((A u − y)²) with A from all nnz incl. the diagonal (model.py:157-167)."""
    r = residual_vector(u, batch)
    return torch.mean(r ** 2)


def residual_vector(u: Tensor, batch) -> Tensor:
    row, col = batch.edge_index[0], batch.edge_index[1]
    Au = torch.zeros_like(u).index_add(0, row, batch.a_ij.reshape(-1, 1) * u[col])
    return Au - batch.y


# ---------------------------------------------------------------------------
# fixed-point solvers
# ---------------------------------------------------------------------------

def _lowrank_apply(U: Tensor, VT: Tensor, x: Tensor) -> Tensor:
    """(−I + U Vᵀ) x   (reference ``matvec``, solver.py:106-114)."""
    if U.nelement() == 0:
        return -x
    return -x + torch.einsum("bijd,bd->bij", U, torch.einsum("bdij,bij->bd", VT, x))


def _lowrank_apply_t(U: Tensor, VT: Tensor, x: Tensor) -> Tensor:
    """xᵀ(−I + U Vᵀ)   (reference ``rmatvec``, solver.py:96-104)."""
    if U.nelement() == 0:
        return -x
    return -x + torch.einsum("bd,bdij->bij", torch.einsum("bij,bijd->bd", x, U), VT)


def broyden(f: Callable[[Tensor], Tensor], x0: Tensor, threshold: int, eps: float = 1e-3,
            keep_trace: bool = True) -> dict:
    """Good-Broyden on g(x)=f(x)−x over the whole batch as one vector, no line search
    (reference ``broyden`` with ``ls=False``, ``stop_mode='rel'``; solver.py:116-207).

    Same tensor layout as the reference — ``Us (1,N,d,thr)``, ``VTs (1,thr,N,d)`` and
    strided slices of them — so that CPU timings of this port are representative.
    """
    x = x0[None]
    _, n_rows, d = x.shape
    g = lambda y: f(y) - y
    gx = g(x[0])[None]
    Us = torch.zeros(1, n_rows, d, threshold, dtype=x0.dtype)
    VTs = torch.zeros(1, threshold, n_rows, d, dtype=x0.dtype)
    step_dir = gx                                             # −(−I)·gx
    protect = 1e3 * d                                         # solver.py:140 ('rel' mode)
    rel_tr: List[float] = []
    abs_tr: List[float] = []
    best = {"rel": 1e8, "abs": 1e8}
    best_step = {"rel": 0, "abs": 0}
    best_x = x[0]
    n = 0
    prot_break = False
    trace = [x[0]]
    while n < threshold:
        x_new = x + step_dir                                  # line_search(on=False): s = 1 (solver.py:85-94)
        g_new = g(x_new[0])[None]
        dx, dg = x_new - x, g_new - gx
        x, gx = x_new, g_new
        if keep_trace:
            trace.append(x[0])
        n += 1
        a = torch.norm(gx).item()
        r = a / (torch.norm(gx + x).item() + 1e-9)
        abs_tr.append(a)
        rel_tr.append(r)
        if r < best["rel"]:
            best_x = x[0].clone().detach()
            best["rel"], best_step["rel"] = r, n
        if a < best["abs"]:
            best["abs"], best_step["abs"] = a, n
        if r < eps:
            break
        if r < 3 * eps and n > 30 and np.max(rel_tr[-30:]) / np.min(rel_tr[-30:]) < 1.3:
            break
        if r > rel_tr[0] * protect:
            prot_break = True
            break
        pU, pV = Us[:, :, :, :n - 1], VTs[:, :n - 1]
        vT = _lowrank_apply_t(pU, pV, dx)
        u = (dx - _lowrank_apply(pU, pV, dg)) / torch.einsum("bij,bij->b", vT, dg)[:, None, None]
        vT[vT != vT] = 0
        u[u != u] = 0
        VTs[:, n - 1] = vT
        Us[:, :, :, n - 1] = u
        step_dir = -_lowrank_apply(Us[:, :, :, :n], VTs[:, :n], gx)
    pad = threshold + 1 - len(rel_tr)
    rel_tr += [best["rel"]] * pad
    abs_tr += [best["abs"]] * pad
    return {"result": best_x, "lowest": best["rel"], "nstep": best_step["rel"], "prot_break": prot_break,
            "abs_trace": abs_tr, "rel_trace": rel_tr, "xest_trace": trace, "eps": eps, "threshold": threshold}


def anderson(f: Callable[[Tensor], Tensor], x0: Tensor, m: int = 2, lam: float = 1e-4, threshold: int = 50,
             eps: float = 1e-3, beta: float = 1.0) -> dict:
    """Anderson acceleration, ``stop_mode='rel'`` (reference ``anderson``, solver.py:215-293)."""
    shape = x0.shape
    nd = x0.numel()
    X = torch.zeros(1, m, nd, dtype=x0.dtype)
    Fm = torch.zeros(1, m, nd, dtype=x0.dtype)
    X[:, 0] = x0.reshape(1, -1)
    Fm[:, 0] = f(x0).reshape(1, -1)
    X[:, 1] = Fm[:, 0]
    Fm[:, 1] = f(Fm[:, 0].reshape(shape)).reshape(1, -1)
    H = torch.zeros(1, m + 1, m + 1, dtype=x0.dtype)
    H[:, 0, 1:] = H[:, 1:, 0] = 1
    y = torch.zeros(1, m + 1, 1, dtype=x0.dtype)
    y[:, 0] = 1
    rel_tr, abs_tr = [], []
    best = {"rel": 1e8, "abs": 1e8}
    best_step = {"rel": 0, "abs": 0}
    best_x = None
    trace = [x0]
    for k in range(2, threshold):
        n = min(k, m)
        G = Fm[:, :n] - X[:, :n]
        H[:, 1:n + 1, 1:n + 1] = torch.bmm(G, G.transpose(1, 2)) + lam * torch.eye(n, dtype=x0.dtype)[None]
        alpha = torch.linalg.solve(H[:, :n + 1, :n + 1], y[:, :n + 1])[:, 1:n + 1, 0]
        X[:, k % m] = beta * (alpha[:, None] @ Fm[:, :n])[:, 0] + (1 - beta) * (alpha[:, None] @ X[:, :n])[:, 0]
        Fm[:, k % m] = f(X[:, k % m].reshape(shape)).reshape(1, -1)
        gx = Fm[:, k % m] - X[:, k % m]
        a = gx.norm().item()
        r = a / (1e-5 + Fm[:, k % m].norm().item())
        abs_tr.append(a)
        rel_tr.append(r)
        if r < best["rel"]:
            best_x = X[:, k % m].reshape(shape).clone().detach()
            best["rel"], best_step["rel"] = r, k
        if a < best["abs"]:
            best["abs"], best_step["abs"] = a, k
        trace.append(best_x)
        if rel_tr[-1] < eps:
            pad = threshold - 1 - k
            rel_tr += [best["rel"]] * pad
            abs_tr += [best["abs"]] * pad
            break
    return {"result": best_x, "lowest": best["rel"], "nstep": best_step["rel"], "prot_break": False,
            "abs_trace": abs_tr, "rel_trace": rel_tr, "xest_trace": trace, "eps": eps, "threshold": threshold}


def forward_iteration(f: Callable[[Tensor], Tensor], z0: Tensor, eps: float = 1e-5, threshold: int = 50) -> dict:
    """Picard iteration (reference ``forward_iteration``, solver.py:301-341)."""
    trace = [z0]
    z_prev, z = z0, f(z0)
    abs_tr = [torch.linalg.norm(z_prev - z)]
    rel_tr = [abs_tr[-1] / torch.linalg.norm(z)]
    trace.append(z)
    it = 0
    while rel_tr[-1] > eps and it < threshold:
        z_prev, z = z, f(z)
        it += 1
        abs_tr.append(torch.linalg.norm(z_prev - z))
        rel_tr.append(abs_tr[-1] / torch.linalg.norm(z))
        trace.append(z)
    return {"result": z, "lowest": rel_tr[-1], "abs_trace": abs_tr, "rel_trace": rel_tr,
            "xest_trace": trace, "nstep": it, "eps": eps, "threshold": threshold}


# ---------------------------------------------------------------------------
# DEQ wrapper and model-level entry points
# ---------------------------------------------------------------------------

def jac_loss(f_out: Tensor, z: Tensor, v: Tensor, create_graph: bool = True) -> Tensor:
    """Hutchinson ‖Jᵀv‖²/numel with ``vecs=1`` and the probe vector supplied (model.py:416-435)."""
    vJ = torch.autograd.grad(f_out, z, v, retain_graph=True, create_graph=create_graph)[0]
    return vJ.norm() ** 2 / 1 / np.prod(z.shape)


def power_method(f_out: Tensor, z: Tensor, evector: Tensor, n_iters: int = 150):
    """Power iteration on Jᵀ (model.py:437-452) with the start vector supplied."""
    for i in range(n_iters):
        vTJ = torch.autograd.grad(f_out, z, evector, retain_graph=(i < n_iters - 1))[0]
        evalue = (vTJ * evector).reshape(1, -1).sum(1, keepdim=True) / (evector * evector).reshape(1, -1).sum(1, keepdim=True)
        evector = (vTJ.reshape(1, -1) / vTJ.reshape(1, -1).norm(dim=1, keepdim=True)).reshape_as(z)
    return evector, torch.abs(evalue)


def inference(P: Params, batch, fw_thres: int, fw_tol: float, mixed: bool = False, solver=broyden) -> dict:
    """``ModelDEQDSS.inference`` (dirichlet/psignn/model.py:99-107) + solver dict."""
    f = f_mixed if mixed else f_dirichlet
    with torch.no_grad():
        h0 = encoder(P, batch.x)
        out = solver(lambda H: f(P, H, h0, batch), h0, threshold=fw_thres, eps=fw_tol)
        out["u"] = decoder(P, out["result"])
    return out


def training_forward_backward(P: Params, batch, fw_thres: int, fw_tol: float, bw_thres: int, bw_tol: float,
                              v: Tensor, jac_weight: float = 1.0, mixed: bool = False, solver=broyden) -> dict:
    """One ``ModelDEQDSS.forward`` + ``loss.backward()`` (dirichlet/psignn/model.py:58-97,185-243;
    step recipe dirichlet/psignn/training_class.py:147-160).  ``v`` is the Hutchinson probe
    (the reference draws it with ``torch.randn`` at model.py:431).  Returns losses, solver
    statistics and parameter gradients keyed like ``P``.
    """
    f = f_mixed if mixed else f_dirichlet
    P = {k: t.detach().clone().requires_grad_(True) for k, t in P.items()}
    h0 = encoder(P, batch.x)
    with torch.no_grad():
        out_fw = solver(lambda H: f(P, H, h0, batch), h0, threshold=fw_thres, eps=fw_tol)
    h_star = out_fw["result"].detach().clone().requires_grad_(True)
    new_h = f(P, h_star, h0, batch)
    jl = jac_loss(new_h, h_star, v)
    stats = {}

    def hook(grad):
        handle[0].remove()                                    # avoid recursion (model.py:211-212)
        out_bw = solver(lambda y: torch.autograd.grad(new_h, h_star, y, retain_graph=True)[0] + grad,
                        torch.zeros_like(grad), threshold=bw_thres, eps=bw_tol)
        stats["bw"] = out_bw
        return out_bw["result"]

    handle = [new_h.register_hook(hook)]
    u = decoder(P, new_h)
    res = residual_loss(u, batch)
    u_d, h_d = u.detach(), new_h.detach()
    enc_loss = F.mse_loss(encoder(P, u_d), h_d)
    ae_loss = F.mse_loss(decoder(P, encoder(P, u_d).detach()), u_d)
    loss = res + jac_weight * jl + enc_loss + ae_loss
    loss.backward()
    grads = {k: (t.grad if t.grad is not None else torch.zeros_like(t)) for k, t in P.items()}
    return {"u": u.detach(), "h_star": h_star.detach(), "loss": loss.detach(), "residual_loss": res.detach(),
            "jacobian_loss": jl.detach(), "encoder_loss": enc_loss.detach(), "autoencoder_loss": ae_loss.detach(),
            "fw": out_fw, "bw": stats.get("bw"), "grads": grads}
