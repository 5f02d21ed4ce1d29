"""psi_gnn_b200 — B200-native implicit message-passing solve of PSI-GNN (see README.md).

Drop-in modules:  psi_gnn_b200.dirichlet.psignn.model / .utilities.solver,  psi_gnn_b200.mixed.psignn.model / .utilities.solver,
psi_gnn_b200.dirichlet.dss.model,  psi_gnn_b200.dirichlet.dsgps.model.   The CUDA extension is loaded lazily by ``_native.load()``
and there is no fallback: without it (or on CPU tensors) every entry point raises ``RuntimeError``.
"""
__version__ = "0.1.0"
