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


# ---- baselines: backward of one unrolled DSS / DSGPS layer, the algebra of csrc/baseline_bwd.cuh ----------------------------------
def baseline_layer_backward(kind, blob, ei, attr, tags, prb, nrm, h, y):
    """(h̄, θ̄ flat in packed-block layout) of one unrolled baseline layer — node record + (dst, ty, tx) table of
    psi_gnn_b200/weights.py, exactly the quantities ``bl_node`` / ``k_bl_gather`` / ``k_bl_pass`` form, vectorised in torch.
    kind: 2 DSS, 3 DSGPS, 4 mixed DSGPS.  ``blob``: packed block (weights.pack_dss / pack_dsgps) in the dtype of ``h``."""
    from psi_gnn_b200 import weights as W
    PG, OFF, D = W.PG, W.OFFSETS, 10
    dss, mixed = kind == 2, kind == 4
    ATTR, PRB = (1 if dss else 3), (2 if kind == 3 else 3)

    def F(name, rows=None, width=None, cols=None):
        o = OFF[name]
        if rows is None:
            return blob[o:o + cols]
        return blob[o:o + rows * width].view(rows, width)[:, :cols if cols is not None else width]

    n = h.size(0)
    keep = ei[0] != ei[1]
    row, col, a = ei[0][keep], ei[1][keep], attr[keep].reshape(int(keep.sum()), -1)[:, :ATTR]
    if dss:
        cls = torch.zeros(n, dtype=torch.long)
    elif mixed:
        cls = torch.where(tags[:, 1] == 1, 1, torch.where(tags[:, 2] == 1, 2, 0))
    else:
        cls = (tags.reshape(-1) == 1).long()
    rec = torch.zeros(n, PG["REC"], dtype=h.dtype)
    rec[:, PG["ONE"]] = 1
    rec[:, PG["C"]:PG["C"] + D] = h
    rec[:, PG["CN"]:PG["CN"] + D] = h

    def edge(slot, dst, src, active):
        W1i, W1j, W1a = F(slot + ".W1i", D, D), F(slot + ".W1j", D, D), F(slot + ".W1a", D, 3, ATTR)
        z = (h @ W1i.T + F(slot + ".b1", cols=D))[dst] + (h @ W1j.T)[src] + a @ W1a.T
        act = active[dst][:, None].to(h.dtype)
        on = (z > 0).to(h.dtype) * act
        S = torch.zeros(n, D, dtype=h.dtype).index_add(0, dst, torch.relu(z) * act)
        cnt = torch.zeros(n, D, dtype=h.dtype).index_add(0, dst, on)
        deg = torch.zeros(n, dtype=h.dtype).index_add(0, dst, act[:, 0])
        Aat = torch.zeros(n, D, 3, dtype=h.dtype)
        Aat[:, :, :ATTR] = torch.zeros(n, D, ATTR, dtype=h.dtype).index_add(0, dst, on[:, :, None] * a[:, None, :])
        mp = S @ F(slot + ".W2", D, D).T + deg[:, None] * F(slot + ".b2", cols=D)
        return on, S, cnt, deg, Aat, mp

    inter, neu = cls == 0, cls == 2
    onT, ST, cT, dT, AT, mT = edge("to", col, row, inter)
    onF, SF, cF, dF, AF, mF = edge("from", row, col, inter)
    Dl = torch.zeros(n, D, dtype=h.dtype)
    c = torch.cat([h, mT, mF, prb[:, :PRB]], 1)
    width = 30 + PRB
    rec[:, PG["C"] + 10:PG["C"] + 30] = torch.cat([mT, mF], 1) * inter[:, None]
    rec[:, PG["C"] + 30:PG["C"] + 30 + PRB] = prb[:, :PRB] * inter[:, None]
    rec[:, PG["DEG"]] = dT
    rec[:, PG["DEG"] + 1] = dF
    yi = y * inter[:, None]
    if dss:
        alpha = blob[OFF["dss_alpha"]]
        pre = c @ F("up_W1", D, 33, width).T + F("up_b1", cols=D)
        hid = torch.relu(pre)
        mb = alpha * yi
        tb = (mb @ F("up_W2", D, D)) * (pre > 0)
        cb = tb @ F("up_W1", D, 33, width)
        Dl = yi + cb[:, :D]
        mTb, mFb = cb[:, D:2 * D], cb[:, 2 * D:3 * D]
        rec[:, PG["MB"]:PG["MB"] + D] = mb
        rec[:, PG["HID"]:PG["HID"] + D] = hid * inter[:, None]
        rec[:, PG["TB"]:PG["TB"] + D] = tb
    else:
        gz, gr, gc = F("gz_W", D, 33, width), F("gr_W", D, 33, width), F("gc_W", D, 33, width)
        zk = torch.sigmoid(c @ gz.T + F("gz_b", cols=D))
        rk = torch.sigmoid(c @ gr.T + F("gr_b", cols=D))
        c2 = torch.cat([rk * h, c[:, D:]], 1)
        th = torch.tanh(c2 @ gc.T + F("gc_b", cols=D))
        eb = yi * zk * (1 - th * th)
        ab = yi * th * zk * (1 - zk)
        c2b = eb @ gc
        bb = c2b[:, :D] * h * rk * (1 - rk)
        cb = ab @ gz + bb @ gr
        Dl = yi + c2b[:, :D] * rk * inter[:, None] + cb[:, :D]
        mTb, mFb = c2b[:, D:2 * D] + cb[:, D:2 * D], c2b[:, 2 * D:3 * D] + cb[:, 2 * D:3 * D]
        rec[:, PG["YB"]:PG["YB"] + D] = eb
        rec[:, PG["TB"]:PG["TB"] + D] = ab
        rec[:, PG["MB"]:PG["MB"] + D] = bb
        rec[:, PG["RHAT"]:PG["RHAT"] + D] = rk * h * inter[:, None]
    Sb = {}

    def tail(slot, w, mpb, S, cnt, Aat):
        nonlocal Dl
        eb_ = PG["EDGE"] + 70 * w
        Sbar = mpb @ F(slot + ".W2", D, D)
        zs = Sbar * cnt
        Dl = Dl + zs @ F(slot + ".W1i", D, D)
        rec[:, eb_:eb_ + 10] = mpb
        rec[:, eb_ + 10:eb_ + 20] = S
        rec[:, eb_ + 20:eb_ + 30] = zs
        rec[:, eb_ + 30:eb_ + 40] = Sbar
        rec[:, eb_ + 40:eb_ + 70] = Aat.reshape(n, 30)
        Sb[slot] = Sbar

    tail("to", 0, mTb, ST, cT, AT)
    tail("from", 1, mFb, SF, cF, AF)
    accN = torch.zeros(n, D, dtype=h.dtype)
    if mixed:
        onN, SN, cN_, dN, AN, mN = edge("neu", row, col, neu)
        yn = y * neu[:, None]
        cn = torch.cat([h, mN, prb[:, :3], nrm], 1)
        pre = cn @ F("un_W1", D, 25).T + F("un_b1", cols=D)
        tbn = (yn @ F("un_W2", D, D)) * (pre > 0)
        cnb = tbn @ F("un_W1", D, 25)
        Dl = Dl + cnb[:, :D]
        rec[:, PG["CN"] + 10:PG["CN"] + 20] = mN * neu[:, None]
        rec[:, PG["CN"] + 20:PG["CN"] + 25] = torch.cat([prb[:, :3], nrm], 1) * neu[:, None]
        rec[:, PG["MBN"]:PG["MBN"] + D] = yn
        rec[:, PG["HIDN"]:PG["HIDN"] + D] = torch.relu(pre) * neu[:, None]
        rec[:, PG["TBN"]:PG["TBN"] + D] = tbn
        rec[:, PG["DEG"] + 2] = dN
        tail("neu", 2, cnb[:, D:2 * D], SN, cN_, AN)
        accN = torch.zeros(n, D, dtype=h.dtype).index_add(0, col, onN * Sb["neu"][row])
    # the node as SOURCE: masks of its outgoing messages times S̄ of their destinations
    accT = torch.zeros(n, D, dtype=h.dtype).index_add(0, row, onT * Sb["to"][col])
    accF = torch.zeros(n, D, dtype=h.dtype).index_add(0, col, onF * Sb["from"][row])
    rec[:, PG["ACC"]:PG["ACC"] + 30] = torch.cat([accT, accF, accN], 1)
    hbar = Dl + accT @ F("to.W1j", D, D) + accF @ F("from.W1j", D, D)
    if mixed:
        hbar = hbar + accN @ F("neu.W1j", D, D)
    dst, ty, tx = (torch.tensor(t) for t in W.baseline_grad_table(kind))
    flat = torch.zeros(W.TOTAL_FLOATS, dtype=h.dtype)
    flat[dst] = (rec[:, ty] * rec[:, tx]).sum(0)
    return hbar, flat
