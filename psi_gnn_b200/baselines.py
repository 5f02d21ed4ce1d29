"""DSS and DSGPS baselines on the shared fused layer kernel (``psi_layer_forward`` kinds 2 and 3).

``DeepStatisticalSolver`` (reference dirichlet/dss/model.py:29-127): k unrolled layers with per-layer weights,
``H ← H + α·Ψ_k(cat[H, ΣΦ→, ΣΦ←, b'])``, edge feature = normalised stiffness coefficient, no LayerNorm, no clamp.
``ModelDSGPS`` (reference dirichlet/dsgps/model.py:27-163): k steps of one GRU-gated recurrent layer + autoencoder.

Class names, constructor signatures and ``state_dict`` keys are the reference's, so its checkpoints load; ``inference`` runs
entirely on the extension (one weight-block upload + one layer launch per unrolled step).  The unrolled *training*
forward/backward of the baselines is outside the accelerated path ("layer kernel reuse only").
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _native as N
from . import weights as W
from .graph import graph_of
from .model import MLP, Autoencoder, Phi_from, Phi_to, initialize_weights_xavier


class Psi(nn.Module):
    def __init__(self, hidden_channels=None, activation=None):
        super().__init__()
        self.mlp = MLP(hidden_channels, activation)

    def forward(self, x):
        return self.mlp(x)


class DecoderDSS(nn.Module):
    def __init__(self, hidden_channels=None, activation=None):
        super().__init__()
        self.mlp = MLP(hidden_channels, activation)

    def forward(self, x):
        return self.mlp(x)


class MLPActivation(nn.Module):
    """Linear → activation (reference dirichlet/dsgps/model.py ``MLPActivation``); keys ``mlp.0.{weight,bias}``."""

    def __init__(self, hidden_channels=None, activation=None):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(hidden_channels[0], hidden_channels[1]), activation).apply(initialize_weights_xavier)

    def forward(self, x):
        return self.mlp(x)


def _decode(h):
    u = torch.empty(h.shape[0], 1, dtype=torch.float32, device=h.device)
    with torch.cuda.device(h.device):
        N.check(N.load().psi_decode(h.shape[0], N.ptr(h), N.ptr(u), N.stream_ptr()), "psi_decode")
    return u


class DeepStatisticalSolver(nn.Module):
    """config keys: latent_dim, k, alpha, gamma (reference dirichlet/dss/main.py)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        d, k = config["latent_dim"], config["k"]
        self.phi_to_list = nn.ModuleList([Phi_to([2 * d + 1, d, d], nn.ReLU()) for _ in range(k)])
        self.phi_from_list = nn.ModuleList([Phi_from([2 * d + 1, d, d], nn.ReLU()) for _ in range(k)])
        self.psi_list = nn.ModuleList([Psi([3 * d + 3, d, d], nn.ReLU()) for _ in range(k)])
        self.decoder_list = nn.ModuleList([DecoderDSS([d, d, 1], nn.ReLU()) for _ in range(k)])
        self.mse_loss = nn.MSELoss()
        self._blobs = (None, None, None)

    def _packed(self, device):
        P = W.named_tensors(self)
        key = (W.version_key(P), str(device))
        if self._blobs[0] != key:
            with torch.no_grad():
                self._blobs = (key, torch.stack([W.pack_dss(P, k, self.config["alpha"], device) for k in range(self.config["k"])]).contiguous(),
                               W.next_serial())
        return self._blobs[2], self._blobs[1]

    def inference(self, batch):
        if not batch.edge_index.is_cuda:
            raise RuntimeError("psi_gnn_b200: CUDA tensors required — there is no CPU path")
        if self.config["latent_dim"] != W.D:
            raise NotImplementedError("psi_gnn_b200: the fused kernel is built for latent_dim=10")
        g = graph_of(batch, N.KIND_DSS)
        dev = batch.edge_index.device
        serial, blobs = self._packed(dev)
        k = self.config["k"]
        H = torch.zeros(g.num_nodes, W.D, dtype=torch.float32, device=dev)
        work, out = torch.empty_like(H), torch.empty_like(H)
        with torch.cuda.device(dev):
            N.check(N.load().psi_layers_unrolled(g.handle, N.KIND_DSS, N.ptr(blobs), k, k, N.ptr(H), None, N.ptr(work), N.ptr(out),
                                                 N.stream_ptr()), "psi_layers_unrolled")
        W.mark_resident(dev, (serial, k - 1))     # the last layer's block (with Decoder_{k-1}) is what the constant bank now holds
        return _decode(out)

    def forward(self, batch):
        raise NotImplementedError("psi_gnn_b200: the unrolled DSS training forward is outside the accelerated path; "
                                  "use inference() (dirichlet/dss/model.py:106-127)")

    def residual_loss(self, U, edge_index, a_ij, y):
        """flux-form residual of the reference (dirichlet/dss/model.py:129-148) — plain torch ops, monitoring only"""
        frm, to = edge_index
        p1 = (1 - y[:, 1:2]) * (-y[:, 0:1]) + y[:, 1:2] * (U - y[:, 2:3])
        flux = torch.zeros_like(U).index_add(0, frm, a_ij.reshape(-1, 1) * (U[to] - U[frm]))
        return torch.mean((p1 + flux) ** 2)


class ModelDSGPS(nn.Module):
    """config keys: latent_dim, k, alpha, gamma (reference dirichlet/dsgps/main.py)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        d = config["latent_dim"]
        self.laynorm = nn.LayerNorm(d)
        self.phi_to = Phi_to([2 * d + 3, d, d], nn.ReLU())
        self.phi_from = Phi_from([2 * d + 3, d, d], nn.ReLU())
        self.z_k = MLPActivation([3 * d + 2, d], nn.Sigmoid())
        self.r_k = MLPActivation([3 * d + 2, d], nn.Sigmoid())
        self.correction = MLPActivation([3 * d + 2, d], nn.Tanh())
        self.autoencoder = Autoencoder([1, d, d], nn.ReLU())
        self.mse_loss = nn.MSELoss()
        self._blob = (None, None, None)

    def inference(self, batch, k=None):
        if not batch.edge_index.is_cuda:
            raise RuntimeError("psi_gnn_b200: CUDA tensors required — there is no CPU path")
        if self.config["latent_dim"] != W.D:
            raise NotImplementedError("psi_gnn_b200: the fused kernel is built for latent_dim=10")
        g = graph_of(batch, N.KIND_DSGPS)
        dev = batch.edge_index.device
        P = W.named_tensors(self)
        key = (W.version_key(P), str(dev))
        if self._blob[0] != key:
            with torch.no_grad():
                self._blob = (key, W.pack_dsgps(P, dev), W.next_serial())
        W.upload(self._blob[1], self._blob[2])
        x = N.f32(batch.x.reshape(-1))
        H0 = torch.empty(x.numel(), W.D, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            N.check(N.load().psi_encode(x.numel(), N.ptr(x), N.ptr(H0), N.stream_ptr()), "psi_encode")
        steps = self.config["k"] if k is None else k
        if steps < 1:
            return _decode(H0)
        work, out = torch.empty_like(H0), torch.empty_like(H0)
        with torch.cuda.device(dev):
            N.check(N.load().psi_layers_unrolled(g.handle, N.KIND_DSGPS, N.ptr(self._blob[1]), 1, steps, N.ptr(H0), N.ptr(H0), N.ptr(work),
                                                 N.ptr(out), N.stream_ptr()), "psi_layers_unrolled")
        return _decode(out)

    def forward(self, batch):
        raise NotImplementedError("psi_gnn_b200: the unrolled DSGPS training forward is outside the accelerated path; "
                                  "use inference() (dirichlet/dsgps/model.py:133-163)")
