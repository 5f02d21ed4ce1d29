"""Torch prototypes of the *restructured* arithmetic the CUDA kernels use.

The kernels do not evaluate the edge MLP the way the reference writes it
(``MLP(cat[h_i,h_j,a_e])`` per edge, model.py:346-350).  They use three exact
algebraic identities, which change rounding order only:

* first edge layer split by input block: ``W1·[h_i;h_j;a] = W1i·h_i + W1j·h_j + W1a·a``
  with ``W1i·h_i + b1`` evaluated once per destination node;
* second edge layer after the sum: ``Σ_e (W2·relu(z_e)+b2) = W2·Σ_e relu(z_e) + deg·b2``;
* the VJP scatters become gathers over the *other* adjacency list, and the destination-side
  term collapses to a per-node ReLU-activity count.

``tests/test_restructured_math.py`` checks these prototypes against the oracle (CPU), which is
what justifies using them in the kernels; the GPU tests then check the kernels themselves.
"""
import torch


def _split(P, pre, d=10):
    W1 = P[pre + ".mlp.mlp.0.weight"]
    return (W1[:, :d], W1[:, d:2 * d], W1[:, 2 * d:], P[pre + ".mlp.mlp.0.bias"],
            P[pre + ".mlp.mlp.2.weight"], P[pre + ".mlp.mlp.2.bias"])


def edge_phase(P, pre, h, dst, src, attr, n):
    """returns S = Σ relu(z), deg, z (per edge)."""
    W1i, W1j, W1a, b1, W2, b2 = _split(P, pre)
    z = (h @ W1i.T + b1)[dst] + (h @ W1j.T)[src] + attr @ W1a.T
    S = torch.zeros(n, z.size(1), dtype=h.dtype).index_add(0, dst, torch.relu(z))
    deg = torch.zeros(n, dtype=h.dtype).index_add(0, dst, torch.ones_like(dst, dtype=h.dtype))
    return S, deg, z


def f_dirichlet_restructured(P, h, h0, batch, pre="deqdss.f", want_cache=False):
    ei = batch.edge_index
    keep = ei[0] != ei[1]
    row, col, attr = ei[0][keep], ei[1][keep], batch.edge_attr[keep]
    n = h.size(0)
    ST, degT, zT = edge_phase(P, pre + ".phi_to_list.0", h, col, row, attr, n)
    SF, degF, zF = edge_phase(P, pre + ".phi_from_list.0", h, row, col, attr, n)
    W2T, b2T = P[pre + ".phi_to_list.0.mlp.mlp.2.weight"], P[pre + ".phi_to_list.0.mlp.mlp.2.bias"]
    W2F, b2F = P[pre + ".phi_from_list.0.mlp.mlp.2.weight"], P[pre + ".phi_from_list.0.mlp.mlp.2.bias"]
    mpT = ST @ W2T.T + degT[:, None] * b2T
    mpF = SF @ W2F.T + degF[:, None] * b2F
    c = torch.cat([h, mpT, mpF, batch.prb_data], 1)
    alpha = torch.sigmoid(c @ P[pre + ".alpha.0.weight"].T + P[pre + ".alpha.0.bias"])
    pre_u = c @ P[pre + ".update_list.0.mlp.0.weight"].T + P[pre + ".update_list.0.mlp.0.bias"]
    m = torch.relu(pre_u) @ P[pre + ".update_list.0.mlp.2.weight"].T + P[pre + ".update_list.0.mlp.2.bias"]
    r = h + alpha * m
    mu = r.mean(1, keepdim=True)
    var = ((r - mu) ** 2).mean(1, keepdim=True)
    rstd = torch.rsqrt(var + 1e-5)
    rhat = (r - mu) * rstd
    out = rhat * P[pre + ".laynorm.weight"] + P[pre + ".laynorm.bias"]
    dmask = batch.tags.reshape(-1) == 1
    out = torch.where(dmask[:, None], h0, out)
    if want_cache:
        return out, dict(row=row, col=col, zT=zT, zF=zF, alpha=alpha, pre_u=pre_u, m=m, rstd=rstd, rhat=rhat, dmask=dmask)
    return out


def vjp_dirichlet_restructured(P, hstar, h0, batch, y, pre="deqdss.f"):
    """(∂f/∂h at hstar)ᵀ y in the two-phase gather form used by the CUDA VJP kernels."""
    _, C = f_dirichlet_restructured(P, hstar, h0, batch, pre, want_cache=True)
    n, d = hstar.shape
    row, col = C["row"], C["col"]
    live = (~C["dmask"])[:, None].to(y.dtype)
    # ---- phase A: per node ---------------------------------------------------
    ghat = P[pre + ".laynorm.weight"] * y
    rbar = C["rstd"] * (ghat - ghat.mean(1, keepdim=True) - C["rhat"] * (ghat * C["rhat"]).mean(1, keepdim=True))
    rbar = rbar * live
    abar = (rbar * C["m"]).sum(1, keepdim=True)
    mbar = C["alpha"] * rbar
    sbar = abar * C["alpha"] * (1 - C["alpha"])
    prebar = (mbar @ P[pre + ".update_list.0.mlp.2.weight"]) * (C["pre_u"] > 0)
    cbar = sbar * P[pre + ".alpha.0.weight"] + prebar @ P[pre + ".update_list.0.mlp.0.weight"]
    D = rbar + cbar[:, :d]
    W1iT, W1jT, _, _, W2T, _ = _split(P, pre + ".phi_to_list.0")
    W1iF, W1jF, _, _, W2F, _ = _split(P, pre + ".phi_from_list.0")
    SbT = cbar[:, d:2 * d] @ W2T          # W2ᵀ·m̄p
    SbF = cbar[:, 2 * d:3 * d] @ W2F
    actT = (C["zT"] > 0).to(y.dtype)
    actF = (C["zF"] > 0).to(y.dtype)
    cntT = torch.zeros(n, d, dtype=y.dtype).index_add(0, col, actT)
    cntF = torch.zeros(n, d, dtype=y.dtype).index_add(0, row, actF)
    D = D + (SbT * cntT) @ W1iT + (SbF * cntF) @ W1iF
    # ---- phase B: gather over the other list -----------------------------------
    accT = torch.zeros(n, d, dtype=y.dtype).index_add(0, row, SbT[col] * actT)   # j = row is the source of 'to' edges
    accF = torch.zeros(n, d, dtype=y.dtype).index_add(0, col, SbF[row] * actF)   # j = col is the source of 'from' edges
    return D + accT @ W1jT + accF @ W1jF
