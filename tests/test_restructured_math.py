"""CPU: the algebraic restructurings the CUDA kernels use (tests/restructured_math.py) change rounding order only."""
import pytest
import torch

from conftest import Golden, rel_err
from restructured_math import f_dirichlet_restructured, vjp_dirichlet_restructured


def test_restructured_forward_fp32_and_fp64():
    for name in ("dirichlet_ckpt", "dirichlet_seed0"):
        g = Golden(name)
        P, b = g.params(), g.batch()
        h0 = g.t("h0")
        out = f_dirichlet_restructured(P, g.t("f1"), h0, b)
        assert rel_err(out, g.t("f2")) <= 1e-6
        P64 = {k: v.double() for k, v in P.items()}
        out64 = f_dirichlet_restructured(P64, g.t("f1").double(), h0.double(), b.double())
        assert rel_err(out64, g.t("f2")) <= 1e-6


def test_restructured_vjp():
    for name in ("dirichlet_ckpt", "dirichlet_seed0"):
        g = Golden(name)
        P, b = g.params(), g.batch()
        out = vjp_dirichlet_restructured(P, g.t("f2"), g.t("h0"), b, g.t("vjp_y"))
        assert rel_err(out, g.t("vjp_out")) <= 1e-5


@pytest.mark.parametrize("name", ["dss_ckpt", "dsgps_ckpt", "dsgps_mixed_ckpt"])
def test_baseline_layer_backward_algebra(name):
    """the node-record / table algebra of csrc/baseline_bwd.cuh (h̄ and every parameter gradient of one unrolled DSS / DSGPS / mixed
    DSGPS layer) against autograd through the torch form of the same layer, in fp64; also checks baseline_grad_table and the
    unpack_*_grads inverses of the packers"""
    from conftest import Golden
    from psi_gnn_b200 import weights as W
    import restructured_math as R
    g = Golden(name)
    z = g.z
    cfg = dict(latent_dim=10, k=int(z["cfg.k"]), alpha=float(z["cfg.alpha"]), gamma=0.9)
    b = g.batch()
    torch.manual_seed(0)
    n = b.num_nodes
    h = torch.randn(n, 10, dtype=torch.float64) * 0.5
    y = torch.randn(n, 10, dtype=torch.float64)
    from oracle import psignn_oracle as O
    P = {kk: v.double().requires_grad_() for kk, v in g.params().items()}
    Pd = {kk: v.detach() for kk, v in P.items()}
    hh = h.clone().requires_grad_()
    for attr in ("edge_attr", "prb_data", "unit_normal_vector", "a_ij_norm", "b_prime_norm", "x"):
        if getattr(b, attr, None) is not None:
            setattr(b, attr, getattr(b, attr).double())
    if name.startswith("dss"):
        k = 3
        out = O.dss_layer(P, k, hh, b, cfg["alpha"])
        blob = _pack64(lambda Pf: W.pack_dss(Pf, k, cfg["alpha"], "cpu"), Pd)
        blob[W.OFFSETS["dss_alpha"]] = cfg["alpha"]          # the packer stores the fp32 rounding of the constant
        hbar, flat = R.baseline_layer_backward(2, blob, b.edge_index, b.a_ij_norm, None, b.b_prime_norm, None, h, y)
        grads = W.unpack_dss_grads(flat, k)
    else:
        mixed = "mixed" in name
        h0 = torch.randn(n, 10, dtype=torch.float64)
        if mixed:
            # one step of mixed/dsgps/model.py:76-97 in torch
            ei, attr = O.offdiag(b.edge_index, b.edge_attr)
            to = O.phi(P, "phi_to", hh, ei, attr, True)
            fr = O.phi(P, "phi_from", hh, ei, attr, False)
            ne = O.phi(P, "phi_neumann", hh, ei, attr, False)
            c = torch.cat([hh, to, fr, b.prb_data], 1)
            zg = torch.sigmoid(O._lin(P, "z_k.mlp.0", c))
            rg = torch.sigmoid(O._lin(P, "r_k.mlp.0", c))
            corr = torch.tanh(O._lin(P, "correction.mlp.0", torch.cat([rg * hh, to, fr, b.prb_data], 1)))
            upd = O.mlp2(P, "update_neumann.mlp", torch.cat([hh, ne, b.prb_data, b.unit_normal_vector], 1))
            out = hh + zg * corr
            out = torch.where((b.tags[:, 2] == 1)[:, None], upd, out)
            out = torch.where((b.tags[:, 1] == 1)[:, None], h0, out)
        else:
            out = O.dsgps_layer(P, hh, h0, b)
        blob = _pack64(lambda Pf: W.pack_dsgps(Pf, "cpu"), Pd)
        hbar, flat = R.baseline_layer_backward(4 if mixed else 3, blob, b.edge_index, b.edge_attr, b.tags, b.prb_data,
                                               getattr(b, "unit_normal_vector", None), h, y)
        grads = W.unpack_dsgps_grads(flat, mixed)
    params = P
    used = [kk for kk in grads]
    ref = torch.autograd.grad(out, [hh] + [params[kk] for kk in used], y, allow_unused=True)
    assert float((hbar - ref[0]).norm() / ref[0].norm()) < 1e-12
    for kk, r in zip(used, ref[1:]):
        r = torch.zeros_like(params[kk]) if r is None else r
        assert grads[kk].shape == r.shape, kk
        assert float((grads[kk] - r).norm()) <= 1e-11 * (1 + float(r.norm())), kk


def _pack64(packer, P):
    """a packed block carrying fp64 values: the packers write fp32, so pack the high and the low fp32 halves and add them"""
    hi = {k: v.float() for k, v in P.items()}
    lo = {k: (v - hi[k].double()).float() for k, v in P.items()}
    b_hi, b_lo = packer(hi).double(), packer(lo).double()
    from psi_gnn_b200 import weights as W
    o = W.OFFSETS["dss_alpha"]
    b_lo[o] = 0.0                     # a constant, not a parameter: written by both packs
    return b_hi + b_lo
