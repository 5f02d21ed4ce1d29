"""DSS and DSGPS baselines on the shared fused layer kernel (``psi_layer_forward`` kinds 2 and 3).

``DeepStatisticalSolver`` (reference dirichlet/dss/model.py:29-127): k unrolled layers with per-layer weights,
``H ← H + α·Ψ_k(cat[H, ΣΦ→, ΣΦ←, b'])``, edge feature = normalised stiffness coefficient, no LayerNorm, no clamp.
``ModelDSGPS`` (reference dirichlet/dsgps/model.py:27-163): k steps of one GRU-gated recurrent layer + autoencoder.

Class names, constructor signatures and ``state_dict`` keys are the reference's, so its checkpoints load; ``inference`` runs
entirely on the extension (one weight-block upload + one layer launch per unrolled step).  ``forward`` (the unrolled training
forward with the per-layer losses of the reference) evaluates every layer with the same native kernel and differentiates it with
``psi_layer_backward`` (``_UnrolledLayer``: h̄ and the parameter gradients of one step, deterministic, no atomics); the decoders /
encoder are ``nn.Linear`` under autograd, the residuals use the native SpMV.  ``_step_torch`` keeps the differentiable torch form of a
DSGPS step for user code that needs a graph through the layer (double backward).  ``ModelDSGPSMixed`` is the mixed-boundary variant
(reference mixed/dsgps/model.py).
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import autograd

from . import _native as N
from . import weights as W
from .graph import graph_of
from .model import MLP, Autoencoder, Phi_from, Phi_to, _ResidualLoss, initialize_weights_xavier


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


class _UnrolledLayer(autograd.Function):
    """one unrolled step ``H_{k+1} = layer_k(H_k)`` of a baseline: forward = the fused layer kernel with the step's weight block,
    backward = ``psi_layer_backward`` (Jᵀȳ and (∂f/∂θ)ᵀȳ in one call) + the gradient to ``h0`` (the clamped Dirichlet rows copy it).
    ``owner`` is the model (``_layer_kind``, ``_layer_names``, ``_layer_unpack``), ``block`` the step's (packed weight block, bank key)
    from ``owner._layer_block``, ``params`` the step's parameters in the order of ``owner._layer_names(step)``."""

    @staticmethod
    def forward(ctx, owner, batch, step, block, dmask, h, h0, *params):
        g = graph_of(batch, owner._layer_kind)
        W.upload(*block)
        ctx.owner, ctx.g, ctx.step, ctx.block, ctx.dmask = owner, g, step, block, dmask
        ctx.save_for_backward(h)
        return g.layer_forward(owner._layer_kind, h.detach(), h0.detach() if h0 is not None else None)

    @staticmethod
    def backward(ctx, ybar):
        (h,) = ctx.saved_tensors
        owner = ctx.owner
        W.upload(*ctx.block)                 # the forward's own packed block: another step's block may occupy the constant bank by now
        ybar = ybar.contiguous()
        hbar, flat = ctx.g.layer_backward(owner._layer_kind, h.detach(), ybar)
        grads = owner._layer_unpack(flat, ctx.step)
        h0bar = torch.where(ctx.dmask, ybar, torch.zeros_like(ybar)) if (ctx.dmask is not None and ctx.needs_input_grad[6]) else None
        # views into `flat` (a fresh tensor per call): AccumulateGrad copies them into the parameters' .grad
        return (None, None, None, None, None, hbar, h0bar) + tuple(grads[n] for n in owner._layer_names(ctx.step))


class _Flux(autograd.Function):
    """``Σ_{e=(i→j)} a_e (u_j − u_i)`` per source row on the native SpMV kernel (forward and adjoint): the reference's
    ``scatter_add`` of per-edge fluxes (dirichlet/dss/model.py:137-145) without float atomics"""

    @staticmethod
    def forward(ctx, u, graph):
        ctx.graph, ctx.shape = graph, u.shape
        return graph.flux(u.detach()).view(u.shape)

    @staticmethod
    def backward(ctx, gout):
        return ctx.graph.flux(gout.contiguous(), transpose=True).view(ctx.shape), None


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

    _layer_kind = N.KIND_DSS

    def _layer_block(self, k, device):
        """(packed block of layer k, constant-bank key)"""
        serial, blobs = self._packed(device)
        return blobs[k], (serial, k)

    def _layer_names(self, k):
        cache = self.__dict__.setdefault("_names", {})
        if k not in cache:
            cache[k] = [f"{m}.{k}.mlp.mlp.{i}.{w}" for m in ("phi_to_list", "phi_from_list", "psi_list") for i in (0, 2) for w in ("weight", "bias")]
        return cache[k]

    def _layer_unpack(self, flat, k):
        return W.unpack_dss_grads(flat, k)

    def _step_torch(self, k, h, h0, batch):
        """layer k in differentiable torch form (dirichlet/dss/model.py:83-91) — not used by ``forward`` (which runs ``_UnrolledLayer``);
        kept for user code that needs an autograd graph through the layer"""
        mess_to = self.phi_to_list[k](h, batch.edge_index, batch.a_ij_norm)
        mess_from = self.phi_from_list[k](h, batch.edge_index, batch.a_ij_norm)
        return h + self.config["alpha"] * self.psi_list[k](torch.cat([h, mess_to, mess_from, batch.b_prime_norm], dim=1))

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
        """unrolled training forward (reference dirichlet/dss/model.py:59-104): k layers with per-layer decoders, flux-form residual of
        every intermediate state weighted by gamma^(k-1-update).  Returns (U dict, loss_dic) like the reference."""
        if not batch.edge_index.is_cuda:
            raise RuntimeError("psi_gnn_b200: CUDA tensors required — there is no CPU path")
        cfg = self.config
        H, U, cumul_res, cumul_mse = {}, {}, {}, {}
        total_loss = None
        H['0'] = torch.zeros([batch.num_nodes, cfg["latent_dim"]], dtype=torch.float, device=batch.x.device)
        U['0'] = self.decoder_list[0](H['0']) + batch.x * 0
        cumul_res['0'] = self._residual_native(U['0'], batch)
        cumul_mse['0'] = self.mse_loss(U['0'], batch.x)
        if cfg["latent_dim"] != W.D:
            raise NotImplementedError("psi_gnn_b200: the fused kernel is built for latent_dim=10")
        P = dict(self.named_parameters())
        serial, blobs = self._packed(batch.x.device)        # packed once per forward (the pack cache is keyed by every parameter's version)
        for update in range(cfg["k"]):
            h = H[str(update)]
            H[str(update + 1)] = _UnrolledLayer.apply(self, batch, update, (blobs[update], (serial, update)), None, h, None,
                                                      *[P[n] for n in self._layer_names(update)])
            U[str(update + 1)] = self.decoder_list[update](H[str(update + 1)])
            cumul_res[str(update + 1)] = self._residual_native(U[str(update + 1)], batch)
            cumul_mse[str(update + 1)] = self.mse_loss(U[str(update + 1)], batch.x)
            term = cumul_res[str(update + 1)] * cfg["gamma"] ** (cfg["k"] - update - 1)
            total_loss = term if total_loss is None else total_loss + term
        return U, {"train_loss": total_loss, "residual_loss": cumul_res, "mse_loss": cumul_mse}

    def _residual_native(self, U, batch):
        """the same flux-form residual with the edge sums on the native kernel (``forward`` uses this one: deterministic)"""
        y = batch.b_prime
        p1 = (1 - y[:, 1:2]) * (-y[:, 0:1]) + y[:, 1:2] * (U - y[:, 2:3])
        return torch.mean((p1 + _Flux.apply(U, graph_of(batch, N.KIND_DSS))) ** 2)

    def residual_loss(self, U, edge_index, a_ij, y):
        """flux-form residual of the reference (dirichlet/dss/model.py:129-148) with the reference's signature (torch ops)"""
        frm, to = edge_index
        p1 = (1 - y[:, 1:2]) * (-y[:, 0:1]) + y[:, 1:2] * (U - y[:, 2:3])
        flux = torch.zeros_like(U).index_add(0, frm, a_ij.reshape(-1, 1) * (U[to] - U[frm]))
        return torch.mean((p1 + flux) ** 2)


class ModelDSGPS(nn.Module):
    """config keys: latent_dim, k, alpha, gamma (reference dirichlet/dsgps/main.py)."""
    KIND = N.KIND_DSGPS
    SECOND = 2

    def __init__(self, config):
        super().__init__()
        self.config = config
        d = config["latent_dim"]
        self.laynorm = nn.LayerNorm(d)
        self.phi_to = Phi_to([2 * d + 3, d, d], nn.ReLU())
        self.phi_from = Phi_from([2 * d + 3, d, d], nn.ReLU())
        self.z_k = MLPActivation([3 * d + self.SECOND, d], nn.Sigmoid())
        self.r_k = MLPActivation([3 * d + self.SECOND, d], nn.Sigmoid())
        self.correction = MLPActivation([3 * d + self.SECOND, d], nn.Tanh())
        self.autoencoder = Autoencoder([1, d, d], nn.ReLU())
        self.mse_loss = nn.MSELoss()
        self._blob = (None, None, None)

    def _layer_block(self, step, dev):
        """(packed block of the recurrent step, constant-bank key) — the same block at every step"""
        P = W.named_tensors(self)
        key = (W.version_key(P), str(dev))
        if self._blob[0] != key:
            with torch.no_grad():
                self._blob = (key, W.pack_dsgps(P, dev), W.next_serial())
        return self._blob[1], self._blob[2]

    @property
    def _layer_kind(self):
        return self.KIND

    def _layer_names(self, step):
        if getattr(self, "_names", None) is None:
            self._names = list(W.unpack_dsgps_grads(torch.zeros(W.TOTAL_FLOATS), self.KIND == N.KIND_DSGPS_MIXED))
        return self._names

    def _layer_unpack(self, flat, step):
        return W.unpack_dsgps_grads(flat, self.KIND == N.KIND_DSGPS_MIXED)

    def inference(self, batch, k=None):
        if not batch.edge_index.is_cuda:
            raise RuntimeError("psi_gnn_b200: CUDA tensors required — there is no CPU path")
        if self.config["latent_dim"] != W.D:
            raise NotImplementedError("psi_gnn_b200: the fused kernel is built for latent_dim=10")
        g = graph_of(batch, self.KIND)
        dev = batch.edge_index.device
        W.upload(*self._layer_block(0, dev))
        x = N.f32(batch.x.reshape(-1))
        H0 = torch.empty(x.numel(), W.D, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            N.check(N.load().psi_encode(x.numel(), N.ptr(x), N.ptr(H0), N.stream_ptr()), "psi_encode")
        steps = self.config["k"] if k is None else k
        if steps < 1:
            return _decode(H0)
        work, out = torch.empty_like(H0), torch.empty_like(H0)
        with torch.cuda.device(dev):
            N.check(N.load().psi_layers_unrolled(g.handle, self.KIND, N.ptr(self._blob[1]), 1, steps, N.ptr(H0), N.ptr(H0), N.ptr(work),
                                                 N.ptr(out), N.stream_ptr()), "psi_layers_unrolled")
        return _decode(out)

    def _step_torch(self, step, h, h0, batch):
        """one recurrent step in differentiable torch form (dirichlet/dsgps/model.py:64-78, mixed/dsgps/model.py:76-97) — not used by
        ``forward`` (which runs ``_UnrolledLayer``); kept for user code that needs an autograd graph through the step"""
        dmask, nmask = self._masks(batch)
        mess_to = self.phi_to(h, batch.edge_index, batch.edge_attr)
        mess_from = self.phi_from(h, batch.edge_index, batch.edge_attr)
        c = torch.cat([h, mess_to, mess_from, batch.prb_data], dim=1)
        alpha, reset = self.z_k(c), self.r_k(c)
        corr = self.correction(torch.cat([reset * h, mess_to, mess_from, batch.prb_data], dim=1))
        nxt = h + alpha * corr
        if self.KIND == N.KIND_DSGPS_MIXED:
            mp_neu = self.phi_neumann(h, batch.edge_index, batch.edge_attr)
            upd = self.update_neumann(torch.cat([h, mp_neu, batch.prb_data, batch.unit_normal_vector], dim=1))
            nxt = torch.where(nmask, upd, nxt)
        return torch.where(dmask, h0, nxt)

    def _masks(self, batch):
        """(Dirichlet rows, Neumann rows or None) as [N, 1] boolean masks"""
        t = batch.tags
        mixed = self.KIND == N.KIND_DSGPS_MIXED
        return ((t[:, 1] if mixed else t.reshape(-1)) == 1)[:, None], ((t[:, 2] == 1)[:, None] if mixed else None)

    def residual_loss(self, u, batch):
        """mean((A u − y)²) on the native SpMV kernels (forward and backward), reference dirichlet/dsgps/model.py:165-176"""
        return _ResidualLoss.apply(u, batch.y, graph_of(batch, self.KIND))

    def forward(self, batch):
        """unrolled training forward with the per-step losses of the reference (dirichlet/dsgps/model.py:48-131, mixed/dsgps/
        model.py:48-131): returns (U dict, loss_dic).  The mixed reference computes the encoder / autoencoder losses on detached
        states, the dirichlet one by freezing the decoder / encoder parameters — both are reproduced."""
        if not batch.edge_index.is_cuda:
            raise RuntimeError("psi_gnn_b200: CUDA tensors required — there is no CPU path")
        cfg = self.config
        mixed = self.KIND == N.KIND_DSGPS_MIXED
        dmask, nmask = self._masks(batch)
        index_dirichlet = torch.where(dmask[:, 0])[0]
        H, U = {}, {}
        cumul_res, cumul_mse, cumul_enc, cumul_autoenc, cumul_mse_dirichlet = {}, {}, {}, {}, {}
        total_loss = None
        U['0'] = batch.x
        cumul_res['0'] = self.residual_loss(U['0'], batch)
        cumul_mse['0'] = self.mse_loss(U['0'], batch.sol)
        H['0'] = self.autoencoder.encoder(U['0'])
        enc, dec = self.autoencoder.encoder, self.autoencoder.decoder
        if cfg["latent_dim"] != W.D:
            raise NotImplementedError("psi_gnn_b200: the fused kernel is built for latent_dim=10")
        P = dict(self.named_parameters())
        names = self._layer_names(0)
        block = self._layer_block(0, batch.edge_index.device)
        for update in range(cfg["k"]):
            key = str(update + 1)
            H[key] = _UnrolledLayer.apply(self, batch, update, block, dmask, H[str(update)], H['0'], *[P[n] for n in names])
            U[key] = dec(H[key])
            cumul_res[key] = self.residual_loss(U[key], batch)
            cumul_mse[key] = self.mse_loss(U[key], batch.sol)
            if mixed:
                u_d, h_d = U[key].detach(), H[key].detach()
                cumul_enc[key] = self.mse_loss(enc(u_d), h_d)
                cumul_autoenc[key] = self.mse_loss(dec(enc(u_d).detach()), u_d)
            else:
                for p in dec.parameters():
                    p.requires_grad = False
                cumul_enc[key] = self.mse_loss(self.autoencoder(H[key], sens="latent"), H[key])
                for p in dec.parameters():
                    p.requires_grad = True
                for p in enc.parameters():
                    p.requires_grad = False
                cumul_autoenc[key] = self.mse_loss(self.autoencoder(U[key], sens="physics"), U[key])
                for p in enc.parameters():
                    p.requires_grad = True
            cumul_mse_dirichlet[key] = self.mse_loss(U[key][index_dirichlet, :], batch.sol[index_dirichlet, :])
            term = cumul_res[key] * cfg["gamma"] ** (cfg["k"] - update - 1) + cumul_enc[key] + cumul_autoenc[key]
            total_loss = term if total_loss is None else total_loss + term
        return U, {"train_loss": total_loss, "residual_loss": cumul_res, "encoder_loss": cumul_enc, "autoencoder_loss": cumul_autoenc,
                   "mse_dirichlet": cumul_mse_dirichlet, "mse_loss": cumul_mse}


class ModelDSGPSMixed(ModelDSGPS):
    """mixed Dirichlet/Neumann DSGPS (reference mixed/dsgps/model.py:27-131): 3-column second member and one-hot tags, Neumann rows
    overwritten by ``update_neumann(cat[H, Σphi_neumann, prb, n̂])`` before the Dirichlet clamp.  Native layer kind 4."""
    KIND = N.KIND_DSGPS_MIXED
    SECOND = 3

    def __init__(self, config):
        super().__init__(config)
        d = config["latent_dim"]
        self.phi_neumann = Phi_from([2 * d + 3, d, d], nn.ReLU())
        self.update_neumann = MLP([2 * d + 5, d, d], nn.ReLU())
