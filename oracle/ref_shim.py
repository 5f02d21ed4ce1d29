"""TEST INFRASTRUCTURE ONLY — loads the *unmodified* reference from /root/reference.

The reference's ``model.py`` imports ``torch_geometric`` and ``torch_sparse``
(``dirichlet/psignn/model.py:10,13-15``), neither of which is installed here
(and there is no network).  The three library entry points it uses are
elementary (index gather, sum-scatter, COO SpMV), so this module registers
minimal stand-ins in ``sys.modules`` and then imports the reference's own
files *verbatim from where they lie* — nothing is copied into this repo.

It only works where ``/root/reference`` exists (the build container).  It is
used by ``oracle/make_golden.py`` to produce the fixtures in ``tests/golden``
and by CPU tests that pin ``oracle/psignn_oracle.py`` against the reference.
Nothing in the product package may import this file.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("PSIGNN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dirichlet", "psignn"))


# -- stand-ins for the third-party API surface the reference touches ---------

class _MessagePassing(torch.nn.Module):
    """``torch_geometric.nn.MessagePassing`` as used at model.py:334-368."""

    def __init__(self, aggr="add", flow="source_to_target"):
        super().__init__()
        assert aggr == "add"
        self.flow = flow

    def propagate(self, edge_index, x=None, edge_attr=None):
        i, j = (1, 0) if self.flow == "source_to_target" else (0, 1)
        msg = self.message(x_i=x[edge_index[i]], x_j=x[edge_index[j]], edge_attr=edge_attr)
        out = torch.zeros(x.size(0), msg.size(1), dtype=msg.dtype, device=msg.device)
        return out.index_add(0, edge_index[i], msg)


def _remove_self_loops(edge_index=None, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], (edge_attr[keep] if edge_attr is not None else None)


class _SparseTensor:
    """``torch_sparse.SparseTensor(row=, col=, value=, sparse_sizes=).matmul``."""

    def __init__(self, row=None, col=None, value=None, sparse_sizes=None):
        self.row, self.col, self.value, self.sizes = row, col, value, sparse_sizes

    def matmul(self, u):
        out = torch.zeros(self.sizes[0], u.size(1), dtype=u.dtype, device=u.device)
        return out.index_add(0, self.row, self.value[:, None] * u[self.col])


def _install_shims():
    if "torch_geometric" in sys.modules and getattr(sys.modules["torch_geometric"], "_psi_shim", False):
        return
    tg = types.ModuleType("torch_geometric")
    tg._psi_shim = True
    tgnn = types.ModuleType("torch_geometric.nn")
    tgnn.MessagePassing = _MessagePassing
    tgnn.MLP = object  # placeholder: shadowed by the reference's own ``class MLP``
    tgu = types.ModuleType("torch_geometric.utils")
    tgu.remove_self_loops = _remove_self_loops
    tg.nn, tg.utils = tgnn, tgu
    ts = types.ModuleType("torch_sparse")
    ts.SparseTensor = _SparseTensor
    sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": tgnn,
                        "torch_geometric.utils": tgu, "torch_sparse": ts})


_CACHE = {}


def load_reference(family: str):
    """Import ``<family>/model.py`` and its ``utilities.solver`` from the reference.

    ``family`` in {"dirichlet/psignn", "mixed/psignn", "dirichlet/dss",
    "dirichlet/dsgps", "mixed/dsgps"}.  Returns ``(model_module, solver_module)``.
    The reference modules use bare names (``import model``, ``from utilities
    import ...``), so each family is imported under a private alias with its
    directory temporarily first on ``sys.path``.
    """
    if family in _CACHE:
        return _CACHE[family]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_shims()
    d = os.path.join(REFERENCE_ROOT, family)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k == "model" or k == "utilities" or k.startswith("utilities.")}
    sys.path.insert(0, d)
    rng_state = torch.get_rng_state()
    try:
        # importing prints "Random seed set as 1234" and reseeds torch (model.py:19)
        model = importlib.import_module("model")
        try:
            solver = importlib.import_module("utilities.solver")
        except ModuleNotFoundError:
            solver = None
    finally:
        sys.path.remove(d)
        mods = {k: sys.modules.pop(k) for k in list(sys.modules)
                if k == "model" or k == "utilities" or k.startswith("utilities.")}
        sys.modules.update(saved)
        torch.set_rng_state(rng_state)
    _CACHE[family] = (model, solver, mods)
    return _CACHE[family]


def load_checkpoint(path_rel: str, family: str):
    """``torch.load`` of a shipped checkpoint; needs ``utilities.solver`` importable
    because ``hyperparameters['solver']`` is a pickled function reference."""
    model, solver, mods = load_reference(family)
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        ck = torch.load(os.path.join(REFERENCE_ROOT, path_rel), map_location="cpu", weights_only=False)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return ck


CHECKPOINTS = {
    "dirichlet/psignn": "dirichlet/psignn/results/constant_dataset/ckpt/best_model.pt",
    "mixed/psignn": "mixed/psignn/results/best_model/ckpt/best_model.pt",
    "dirichlet/dss": "dirichlet/dss/results/dss_results/ckpt/best_model.pt",
    "dirichlet/dsgps": "dirichlet/dsgps/results/constant_dataset/30_ite_gamma_0_9/ckpt/best_model.pt",
    "mixed/dsgps": "mixed/dsgps/results/30_ite_lamb_0_gamma_0_9/ckpt/best_model.pt",
}
