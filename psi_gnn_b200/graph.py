"""One-off re-layout of a PyG-style batch into the native destination-sorted lists (csrc/graph.cuh).

Replaces what the reference recomputes on every evaluation of f: ``remove_self_loops``
(dirichlet/psignn/model.py:342,360), ``torch.where(batch.tags == 1)`` (:281) and the
``SparseTensor`` construction (:159-163).  The handle is cached on the batch object.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_int64, c_void_p
from typing import Optional

import torch

from . import _native as N

KIND_NAMES = {N.KIND_DIRICHLET: "dirichlet", N.KIND_MIXED: "mixed", N.KIND_DSS: "dss", N.KIND_DSGPS: "dsgps"}


class NativeGraph:
    """Owns a ``psi_graph_t`` handle (destroyed with the object)."""

    def __init__(self, num_nodes: int, edge_index: torch.Tensor, edge_attr: torch.Tensor, a_ij: Optional[torch.Tensor],
                 tags: Optional[torch.Tensor], prb: Optional[torch.Tensor], normals: Optional[torch.Tensor] = None):
        lib = N.load()
        if edge_index.dtype != torch.int64:
            raise RuntimeError("psi_gnn_b200: edge_index must be int64 (PyG convention)")
        if edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise RuntimeError("psi_gnn_b200: edge_index must have shape [2, nnz]")
        nnz = int(edge_index.shape[1])
        ei = edge_index.contiguous()
        attr = N.f32(edge_attr.reshape(nnz, -1)) if nnz > 0 else torch.zeros(0, 3, dtype=torch.float32, device=ei.device)
        aij = N.f32(a_ij.reshape(-1)) if a_ij is not None else None
        if aij is not None and aij.numel() != nnz:
            raise RuntimeError("psi_gnn_b200: a_ij must have one entry per edge")
        tg = N.f32(tags.reshape(num_nodes, tags.shape[-1] if tags.dim() > 1 else 1)) if tags is not None else None
        pr = N.f32(prb.reshape(num_nodes, prb.shape[-1] if prb.dim() > 1 else 1)) if prb is not None else None
        nr = N.f32(normals.reshape(num_nodes, 2)) if normals is not None else None
        self.device = ei.device
        self.num_nodes = int(num_nodes)
        self.handle = c_void_p()
        with torch.cuda.device(self.device):
            N.check(lib.psi_graph_create(byref(self.handle), self.num_nodes, nnz, N.ptr(ei), N.ptr(attr), int(attr.shape[1]),
                                         N.ptr(aij), N.ptr(tg), int(tg.shape[1]) if tg is not None else 1, N.ptr(pr),
                                         int(pr.shape[1]) if pr is not None else 0, N.ptr(nr), N.stream_ptr()), "psi_graph_create")
        info = (c_int64 * 8)()
        N.check(lib.psi_graph_info(self.handle, info), "psi_graph_info")
        self.num_offdiag, self.nnz, self.num_dirichlet, self.num_neumann = int(info[1]), int(info[2]), int(info[3]), int(info[4])
        self.slots_to, self.slots_from, self.bytes = int(info[5]), int(info[6]), int(info[7])

    def __del__(self):
        try:
            if self.handle:
                N.load().psi_graph_destroy(self.handle)
                self.handle = c_void_p()
        except Exception:
            pass

    # ---- single applications ------------------------------------------------------------------------
    def layer_forward(self, kind: int, h: torch.Tensor, h0: Optional[torch.Tensor]) -> torch.Tensor:
        h = N.f32(h)
        h0c = N.f32(h0) if h0 is not None else None
        out = torch.empty_like(h)
        with torch.cuda.device(self.device):
            N.check(N.load().psi_layer_forward(self.handle, kind, N.ptr(h), N.ptr(h0c), N.ptr(out), N.stream_ptr()), "psi_layer_forward")
        return out

    def vjp_prepare(self, kind: int, hstar: torch.Tensor, h0: Optional[torch.Tensor] = None):
        hs = N.f32(hstar)
        with torch.cuda.device(self.device):
            N.check(N.load().psi_vjp_prepare(self.handle, kind, N.ptr(hs), None, N.stream_ptr()), "psi_vjp_prepare")

    def vjp_apply(self, kind: int, y: torch.Tensor, grad: Optional[torch.Tensor] = None) -> torch.Tensor:
        y = N.f32(y)
        g = N.f32(grad) if grad is not None else None
        out = torch.empty_like(y)
        with torch.cuda.device(self.device):
            N.check(N.load().psi_vjp_apply(self.handle, kind, N.ptr(y), N.ptr(g), N.ptr(out), N.stream_ptr()), "psi_vjp_apply")
        return out

    def param_grad(self, kind: int, hstar: torch.Tensor, ybar: torch.Tensor, want_jty: bool = False):
        """θ̄ = (∂f/∂θ)ᵀ ȳ at the prepared point as a flat vector in packed-block layout (weights.unpack_psignn_grads turns it into
        per-parameter gradients) — native, deterministic; optionally also Jᵀȳ"""
        from . import weights as W
        hs, yb = N.f32(hstar), N.f32(ybar)
        dst, ty, tx = W.grad_table_device(kind == N.KIND_MIXED, self.device)
        out = torch.zeros(W.TOTAL_FLOATS, dtype=torch.float32, device=self.device)
        jty = torch.empty_like(yb) if want_jty else None
        with torch.cuda.device(self.device):
            N.check(N.load().psi_param_grad(self.handle, kind, N.ptr(hs), N.ptr(yb), N.ptr(dst), N.ptr(ty), N.ptr(tx), int(dst.numel()),
                                            N.ptr(out), N.ptr(jty), N.stream_ptr()), "psi_param_grad")
        return (out, jty) if want_jty else out

    def param_grad_tangent(self, kind: int, hstar: torch.Tensor, ybar: torch.Tensor, hdot: torch.Tensor) -> torch.Tensor:
        """d/dε θ̄(H* + ε ḣ; ȳ) at ε = 0 (flat, packed-block layout): with ȳ = v and ḣ = Jᵀv it is ½ ∇θ ‖Jᵀv‖², the double backward of
        the Hutchinson regulariser (model.py:207, :416-435) — native, deterministic"""
        from . import weights as W
        hs, yb, hd = N.f32(hstar), N.f32(ybar), N.f32(hdot)
        dst, ty, tx = W.grad_table_device(kind == N.KIND_MIXED, self.device)
        out = torch.zeros(W.TOTAL_FLOATS, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(N.load().psi_param_grad_tangent(self.handle, kind, N.ptr(hs), N.ptr(yb), N.ptr(hd), N.ptr(dst), N.ptr(ty), N.ptr(tx),
                                                    int(dst.numel()), N.ptr(out), N.stream_ptr()), "psi_param_grad_tangent")
        return out

    def layer_backward(self, kind: int, h: torch.Tensor, ybar: torch.Tensor):
        """(h̄, θ̄) of ONE unrolled baseline layer (kinds DSS / DSGPS / mixed DSGPS) at its own input ``h``: ``h̄ = Jᵀȳ`` and the parameter
        gradient as a flat vector in packed-block layout (weights.unpack_dss_grads / unpack_dsgps_grads) — native, deterministic.
        The layer's weight block must be resident."""
        from . import weights as W
        hs, yb = N.f32(h), N.f32(ybar)
        dst, ty, tx = W.baseline_grad_table_device(kind, self.device)
        out = torch.zeros(W.TOTAL_FLOATS, dtype=torch.float32, device=self.device)
        hbar = torch.zeros_like(hs)
        with torch.cuda.device(self.device):
            N.check(N.load().psi_layer_backward(self.handle, kind, N.ptr(hs), N.ptr(yb), N.ptr(dst), N.ptr(ty), N.ptr(tx), int(dst.numel()),
                                                N.ptr(hbar), N.ptr(out), N.stream_ptr()), "psi_layer_backward")
        return hbar, out

    def residual(self, u: torch.Tensor, y: torch.Tensor, want_vector: bool = False):
        """(mean((A u − y)²), residual vector or None)  — dirichlet/psignn/model.py:157-167."""
        u = N.f32(u.reshape(-1))
        y = N.f32(y.reshape(-1))
        r = torch.empty_like(u) if want_vector else None
        ms = torch.empty(1, dtype=torch.float32, device=u.device)
        with torch.cuda.device(self.device):
            N.check(N.load().psi_residual(self.handle, N.ptr(u), N.ptr(y), N.ptr(r), N.ptr(ms), N.stream_ptr()), "psi_residual")
        return ms[0], r

    def spmv_t(self, v: torch.Tensor) -> torch.Tensor:
        v = N.f32(v.reshape(-1))
        out = torch.empty_like(v)
        with torch.cuda.device(self.device):
            N.check(N.load().psi_spmv_t(self.handle, N.ptr(v), N.ptr(out), N.stream_ptr()), "psi_spmv_t")
        return out

    def flux(self, v: torch.Tensor, transpose: bool = False) -> torch.Tensor:
        """Σ_{e=(i→j)} a_e (v_j − v_i) per source row i (DSS flux-form residual, dirichlet/dss/model.py:137-145), or its adjoint"""
        v = N.f32(v.reshape(-1))
        out = torch.empty_like(v)
        with torch.cuda.device(self.device):
            N.check(N.load().psi_flux(self.handle, N.ptr(v), N.ptr(out), 1 if transpose else 0, N.stream_ptr()), "psi_flux")
        return out

    def solver(self, threshold: int):
        """the solver workspace for this graph's size, taken from a process-wide pool keyed by (device, numel): batches of
        a data loader have a few distinct sizes, and re-allocating GBs of U/V history per batch would dominate the step"""
        from .solver import workspace_for
        return workspace_for(self.num_nodes * 10, threshold, self.device)


def _cache_slot(batch):
    d = getattr(batch, "__dict__", None)
    return d if isinstance(d, dict) else None


# batch attributes whose contents psi_graph_create snapshots (deep copies), per layer kind
_CAPTURED = {
    N.KIND_DIRICHLET: ("edge_index", "edge_attr", "a_ij", "tags", "prb_data"),
    N.KIND_DSGPS: ("edge_index", "edge_attr", "a_ij", "tags", "prb_data"),
    N.KIND_MIXED: ("edge_index", "edge_attr", "a_ij", "tags", "prb_data", "unit_normal_vector"),
    N.KIND_DSGPS_MIXED: ("edge_index", "edge_attr", "a_ij", "tags", "prb_data", "unit_normal_vector"),
    N.KIND_DSS: ("edge_index", "a_ij_norm", "a_ij", "b_prime_norm"),
}


def _stamp(batch, kind: int):
    """identity + in-place version + shape of every tensor the native handle snapshots.  The handle holds deep copies
    (edge records, a_ij, tags, prb, normals), so a batch whose right-hand side, boundary data or normalisation was reassigned
    or modified in place while ``edge_index`` stayed the same must rebuild — otherwise the native solves and the live-tensor
    autograd path would evaluate different functions.  (``t.data`` writes do not bump ``_version``: call :func:`invalidate`.)"""
    out = []
    for name in _CAPTURED[kind]:
        t = getattr(batch, name, None)
        out.append(None if t is None else (t.data_ptr(), t._version, tuple(t.shape), str(t.device)))
    return tuple(out)


def invalidate(batch) -> None:
    """drop every cached native graph of ``batch`` (after modifying captured tensors through ``.data`` or other version-less writes)"""
    slot = _cache_slot(batch)
    if slot is not None:
        for k in [k for k in slot if k.startswith("_psi_graph_") or k.startswith("_psi_offdiag")]:
            del slot[k]


def graph_of(batch, kind: int) -> NativeGraph:
    """The (cached) native graph of a batch for a layer kind.

    PSI-GNN / DSGPS use ``edge_attr [nnz,3]`` and ``prb_data``; DSS uses the 1-column normalised stiffness
    ``a_ij_norm`` and ``b_prime_norm`` on its own diagonal-free edge list (dirichlet/dss/utilities/reader.py).
    """
    slot = _cache_slot(batch)
    key = "_psi_graph_%d" % kind
    ei = batch.edge_index
    stamp = _stamp(batch, kind)
    if slot is not None and key in slot:
        g, old = slot[key]
        if old == stamp:
            return g
    if not ei.is_cuda:
        raise RuntimeError("psi_gnn_b200: the batch must live on a CUDA device (no CPU fallback exists)")
    n = int(batch.num_nodes) if getattr(batch, "num_nodes", None) is not None else int(batch.x.shape[0])
    if kind == N.KIND_DSS:
        g = NativeGraph(n, ei, batch.a_ij_norm, getattr(batch, "a_ij", None), None, batch.b_prime_norm)
    elif kind in (N.KIND_MIXED, N.KIND_DSGPS_MIXED):
        g = NativeGraph(n, ei, batch.edge_attr, batch.a_ij, batch.tags, batch.prb_data, batch.unit_normal_vector)
    else:
        g = NativeGraph(n, ei, batch.edge_attr, batch.a_ij, batch.tags, batch.prb_data)
    part = getattr(batch, "partition", None)
    if part is not None and part.world > 1:
        # one rank's share of a partitioned mesh (psi_gnn_b200/partition.py): bind the halo lists and the communicator
        from . import partition as PT
        if part.comm is None:
            part.comm = PT.Communicator(ei.device)
        PT.attach(g, part, part.comm)
    if slot is not None:
        slot[key] = (g, stamp)
    return g
