"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes) and through the drop-in modules, against
the golden vectors produced by the unmodified reference (tests/golden, oracle/make_golden.py).

Tolerances follow SURVEY §8c: rel L2 ≤ 1e-5 for single applications (f, VJP, residual, encoder/decoder) and for
teacher-forced quasi-Newton steps (against the fp64 evaluation of the reference formulas on the same fp32 inputs, with
the reference's own fp32 deviation as floor); free-running solves are compared on the early trajectory, the stopping
statistics and the converged solution band.
"""
import numpy as np
import pytest
import torch

from conftest import Golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


def band_u(golden):
    """Allowed deviation of a free-running solution: 5 × the distance between two runs of the REFERENCE ITSELF that differ only by a
    permutation of the edge list (stored in the fixture as perm_u), floor 1e-5.  For short solves (random-init fixtures: 17–19 steps)
    the reference is permutation-stable to 2e-7 and the band is the 1e-5 parity bar; for the trained checkpoints (67–191 steps,
    spectral radius 0.99) the reference's own scatter is 4e-4 (dirichlet) to 5e-3 (mixed)."""
    self_dev = rel_err(golden.t("perm_u"), golden.t("u"))
    return max(TOL, 5.0 * self_dev)


def nstep_stable(golden):
    return int(golden["perm_fw_nstep"]) == int(golden["fw_nstep"])



def _kind(g):
    from psi_gnn_b200 import _native as N
    return N.KIND_MIXED if g.mixed else N.KIND_DIRICHLET


def test_extension_loaded():
    from psi_gnn_b200 import _native as N
    assert N.load().psi_version() >= 100


def test_layer_forward_matches_reference(golden):
    m = golden.model(DEV)
    b = golden.batch(DEV)
    h0 = golden.t("h0", DEV)
    with torch.no_grad():
        f1 = m.deqdss.f(h0, h0, b)
        f2 = m.deqdss.f(f1, h0, b)
    assert rel_err(f1, golden.t("f1")) <= TOL
    assert rel_err(f2, golden.t("f2")) <= TOL
    # Dirichlet rows are copied, not computed: bit-exact
    t = b.tags.reshape(b.num_nodes, -1)
    d = (t[:, 0] if t.shape[1] == 1 else t[:, 1]) == 1
    assert torch.equal(f1[d], h0[d])


def test_layer_forward_deterministic(golden):
    m = golden.model(DEV)
    b = golden.batch(DEV)
    h0 = golden.t("h0", DEV)
    with torch.no_grad():
        a = m.deqdss.f(h0, h0, b)
        c = m.deqdss.f(h0, h0, b)
    assert torch.equal(a, c)          # segmented sums, no atomics


def test_encoder_decoder(golden):
    m = golden.model(DEV)
    b = golden.batch(DEV)
    h0 = m._encode_native(b.x)
    assert rel_err(h0, golden.t("h0")) <= TOL
    u = m._decode_native(golden.t("fw_result", DEV))
    assert rel_err(u, golden.t("u")) <= TOL


def test_residual(golden):
    m = golden.model(DEV)
    b = golden.batch(DEV)
    r = m.residual_loss(golden.t("u", DEV), b)
    assert abs(r.item() - float(golden["residual"])) <= 1e-5 * abs(float(golden["residual"]))
    rx = m.residual_loss(b.x, b)
    assert abs(rx.item() - float(golden["residual_x"])) <= 1e-5 * abs(float(golden["residual_x"]))


def test_residual_backward(golden):
    m = golden.model(DEV)
    b = golden.batch(DEV)
    u = golden.t("u", DEV).clone().requires_grad_()
    m.residual_loss(u, b).backward()
    row, col = b.edge_index
    u2 = golden.t("u", DEV).double().requires_grad_()
    Au = torch.zeros_like(u2).index_add(0, row, b.a_ij.double().reshape(-1, 1) * u2[col])
    ((Au - b.y.double()) ** 2).mean().backward()
    assert rel_err(u.grad, u2.grad) <= TOL


def test_vjp_matches_reference(golden):
    from psi_gnn_b200.solver import VjpOperator
    m = golden.model(DEV)
    b = golden.batch(DEV)
    y = golden.t("vjp_y", DEV)
    op = VjpOperator(m.deqdss.f, golden.t("f2", DEV), b, torch.zeros_like(y))
    out = op(y)
    assert rel_err(out, golden.t("vjp_out")) <= TOL
    # + grad term
    g = torch.randn_like(y)
    op2 = VjpOperator(m.deqdss.f, golden.t("f2", DEV), b, g)
    assert rel_err(op2(y) - g, golden.t("vjp_out")) <= 2 * TOL


def test_vjp_is_transpose_of_torch_jvp(golden):
    """⟨Jᵀy, w⟩ = ⟨y, J w⟩ with J w from a central finite difference of the CUDA layer in the direction w (fp32: loose)."""
    from psi_gnn_b200.solver import VjpOperator
    m = golden.model(DEV)
    b = golden.batch(DEV)
    h0, H = golden.t("h0", DEV), golden.t("f2", DEV)
    y = golden.t("vjp_y", DEV)
    w = torch.randn_like(H)
    op = VjpOperator(m.deqdss.f, H, b, torch.zeros_like(y))
    lhs = (op(y).double() * w.double()).sum()
    # torch differentiable path of the drop-in module (autograd) as independent check of the same operator
    Hr = H.clone().requires_grad_()
    with torch.enable_grad():
        out = m.deqdss.f(Hr, h0, b)
    rhs = (torch.autograd.grad(out, Hr, y)[0].double() * w.double()).sum()
    assert abs(lhs - rhs) <= 1e-4 * abs(rhs) + 1e-6


def test_forced_quasi_newton_steps(golden_small):
    """Teacher-forced rank-one updates with the production kernels vs the reference formulas in fp64 on the same inputs."""
    import ctypes
    from psi_gnn_b200 import _native as N
    from psi_gnn_b200.solver import SolverWorkspace
    g = golden_small
    lib = N.load()
    for n in [int(s) for s in g["forced_steps"]]:
        pre = "forced%d_" % n
        x, gx, xn, gn = (g.t(pre + k, DEV) for k in ("x", "gx", "xnew", "gnew"))
        U, V = g.t(pre + "U", DEV).contiguous(), g.t(pre + "V", DEV).contiguous()
        numel = x.numel()
        ws = SolverWorkspace(numel, 64, torch.device(DEV))
        u, v, upd = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        N.check(lib.psi_broyden_forced_step(ws.handle, n, N.ptr(x), N.ptr(gx), N.ptr(xn), N.ptr(gn), N.ptr(U), N.ptr(V),
                                            N.ptr(u), N.ptr(v), N.ptr(upd), N.stream_ptr()), "forced")
        torch.cuda.synchronize()
        for name, got in (("u", u), ("v", v), ("upd", upd)):
            truth = g.t(pre + name + "64")
            ref32 = g.t(pre + name + "32")
            floor = rel_err(ref32, truth)
            if name == "upd":
                # the kernels return δx = (x + update) − x, the reference's own delta_x (solver.py:89-94): its rounding is
                # half an ulp of x per component, i.e. 2⁻²⁴·‖x‖/‖update‖ relative to the update
                floor = max(floor, 2.0 ** -24 * float(xn.double().norm() / truth.norm()))
            e = rel_err(got, truth)
            assert e <= max(TOL, 2 * floor), (n, name, e, floor)
        ws.close()


def test_param_grad_matches_autograd(golden):
    """psi_param_grad — θ̄ = (∂f/∂θ at H)ᵀ ȳ — against torch autograd on the differentiable form of the same layer evaluated in fp64
    (every parameter tensor of the layer, both families, checkpoint and random-init weights); and bit-identical on a second run"""
    from psi_gnn_b200 import weights as W
    from psi_gnn_b200.solver import VjpOperator
    m = golden.model(DEV)
    b = golden.batch(DEV)
    f = m.deqdss.f
    H, y, h0 = golden.t("f2", DEV), golden.t("vjp_y", DEV), golden.t("h0", DEV)
    op = VjpOperator(f, H, b, torch.zeros_like(y))
    flat, jty = op.graph.param_grad(f.kind, H, y, want_jty=True)
    flat2 = op.graph.param_grad(f.kind, H, y)
    assert torch.equal(flat, flat2)                      # fixed reduction order, no atomics
    assert rel_err(jty, golden.t("vjp_out")) <= TOL
    names = ["deqdss.f." + n for n, _ in f.named_parameters()]
    got = W.unpack_psignn_grads(flat, names, golden.mixed)
    assert set(got) == set(names)
    m64 = golden.model(DEV).double()
    b64 = b.double()
    Hr = H.double().requires_grad_()
    out = m64.deqdss.f._forward_torch(Hr, h0.double(), b64)
    ref = torch.autograd.grad(out, list(m64.deqdss.f.parameters()), y.double(), allow_unused=True)
    tot_e, tot_n = 0.0, 0.0
    for (n, p), r in zip(m64.deqdss.f.named_parameters(), ref):
        r = torch.zeros_like(p) if r is None else r
        g_ = got["deqdss.f." + n].double()
        assert g_.shape == r.shape, n
        tot_e += float((g_ - r).norm() ** 2); tot_n += float(r.norm() ** 2)
        assert float((g_ - r).norm()) <= 2e-5 * float(r.norm()) + 1e-7 * (tot_n ** 0.5 + 1e-30), (n, float((g_ - r).norm()), float(r.norm()))
    assert (tot_e / tot_n) ** 0.5 <= TOL


def test_jacobian_loss_double_backward_matches_autograd(golden):
    """psi_param_grad_tangent — the double backward of the Hutchinson regulariser ‖Jᵀv‖²/numel (model.py:207, :416-435) — against torch
    autograd (``create_graph=True`` + backward) on the differentiable form of the same layer in fp64, same probe v: the loss value and
    every parameter gradient, both families, checkpoint and random-init weights; bit-identical on a second run"""
    from psi_gnn_b200 import model as PM
    m = golden.model(DEV)
    b = golden.batch(DEV)
    f = m.deqdss.f
    H, v, h0 = golden.t("f2", DEV), golden.t("vjp_y", DEV), golden.t("h0", DEV)
    params = list(f.parameters())

    def native():
        for p_ in params:
            p_.grad = None
        loss = PM._JacobianLoss.apply(m.deqdss, b, H, v, *params)
        loss.backward()
        return float(loss), {n: (p_.grad.clone() if p_.grad is not None else torch.zeros_like(p_)) for n, p_ in f.named_parameters()}

    l1, g1 = native()
    l2, g2 = native()
    assert l1 == l2 and all(torch.equal(g1[n], g2[n]) for n in g1)
    m64 = golden.model(DEV).double()
    b64 = b.double()
    Hr = H.double().requires_grad_()
    out = m64.deqdss.f._forward_torch(Hr, h0.double(), b64)
    vJ = torch.autograd.grad(out, Hr, v.double(), create_graph=True)[0]
    lref = vJ.norm() ** 2 / vJ.numel()
    ref = torch.autograd.grad(lref, list(m64.deqdss.f.parameters()), allow_unused=True)
    assert abs(l1 - float(lref)) <= 1e-5 * abs(float(lref))
    tot_e, tot_n = 0.0, 0.0
    for (n, p_), r in zip(m64.deqdss.f.named_parameters(), ref):
        r = torch.zeros_like(p_) if r is None else r
        tot_e += float((g1[n].double() - r).norm() ** 2); tot_n += float(r.norm() ** 2)
    for (n, p_), r in zip(m64.deqdss.f.named_parameters(), ref):
        r = torch.zeros_like(p_) if r is None else r
        e = float((g1[n].double() - r).norm())
        assert e <= 5e-5 * float(r.norm()) + 2e-6 * tot_n ** 0.5, (n, e, float(r.norm()), tot_n ** 0.5)
    assert (tot_e / tot_n) ** 0.5 <= 2e-5, (tot_e / tot_n) ** 0.5


def test_forced_anderson_steps(golden_small):
    """Teacher-forced Anderson updates (Gram matrix, bordered solve, mixing) with the production kernels vs the reference's formulas
    (solver.py:250-255) evaluated in fp64 on the same fp32 window: rel L2 ≤ 1e-5 (the reference's own fp32 evaluation is at 2e-8 … 6e-8)"""
    from psi_gnn_b200 import _native as N
    from psi_gnn_b200.solver import SolverWorkspace
    g = golden_small
    lib = N.load()
    m, lam, beta = int(g["andf_m"]), float(g["andf_lam"]), float(g["andf_beta"])
    for k in [int(s) for s in g["andf_steps"]]:
        pre = "andf%d_" % k
        X, F = g.t(pre + "X", DEV).contiguous(), g.t(pre + "F", DEV).contiguous()
        n, numel = X.shape[0], X.shape[1]
        ws = SolverWorkspace(numel, 64, torch.device(DEV))
        xn = torch.empty(numel, device=DEV)
        al = torch.empty(n, device=DEV)
        N.check(lib.psi_anderson_forced_step(ws.handle, m, n, int(g[pre + "slot"]), lam, beta, N.ptr(X), N.ptr(F), N.ptr(xn), N.ptr(al),
                                             N.stream_ptr()), "forced anderson")
        torch.cuda.synchronize()
        floor = rel_err(g.t(pre + "x32"), g.t(pre + "x64"))
        e = rel_err(xn, g.t(pre + "x64"))
        assert e <= max(TOL, 2 * floor), (k, e, floor)
        assert rel_err(al, g.t(pre + "alpha64")) <= 1e-4, (k, al, g[pre + "alpha64"])     # α solves a 4×4 system with cond ~ 1e4
        ws.close()


def test_forward_solve_free_running(golden):
    m = golden.model(DEV)
    b = golden.batch(DEV)
    h0 = golden.t("h0", DEV)
    out = m.deqdss.inference(h0, b, keep_trace=True)
    ref_rel = golden["fw_rel_trace"]
    ref_steps = int(golden["fw_steps_run"])
    # early trajectory: the first iterates agree to fp32 accuracy
    for i, key in ((1, "fw_x1"), (2, "fw_x2"), (3, "fw_x3")):
        assert rel_err(out["xest_trace"][i], golden.t(key)) <= 1e-4
    # the secant updates amplify fp32 reordering noise chaotically after ~10-30 steps (the reference itself
    # moves by O(1) in the residual trace under an edge permutation), so only the first 10 steps are held tight
    k = min(10, ref_steps, out["steps_run"])
    got = np.asarray(out["rel_trace"][:k])
    assert np.all(np.abs(got - ref_rel[:k]) <= 1e-3 * ref_rel[:k] + 1e-9)
    # stopping statistics
    eps = float(golden["cfg.fw_tol"])
    assert out["lowest"] < eps
    assert not out["prot_break"]
    # step counts of converged runs scatter widely between arithmetically equivalent implementations (chaotic secant updates;
    # observed 60–152 for the reference's 80 across kernel revisions that all pass the teacher-forced step test)
    assert out["nstep"] <= 2.5 * int(golden["fw_nstep"]) + 10
    assert len(out["rel_trace"]) == int(golden["cfg.fw_thres"]) + 1
    # the fixed point itself: both are eps-accurate solutions of the same contraction
    # two eps-accurate fixed points of a map with spectral radius ρ ≈ 0.99 differ by up to ~2·eps/(1−ρ) ≈ 3e-3 in h (more in u): band
    u = m._decode_native(out["result"])
    dev = rel_err(u, golden.t("u"))
    print("free-running %s: nstep %d (reference %d, permuted reference %d), u rel diff %.2e (reference vs itself %.2e)" % (
        golden.name, out["nstep"], int(golden["fw_nstep"]), int(golden["perm_fw_nstep"]), dev, rel_err(golden.t("perm_u"), golden.t("u"))))
    assert dev <= band_u(golden)
    if nstep_stable(golden):            # where the reference's own step count is permutation-invariant, ours must equal it
        assert out["nstep"] == int(golden["fw_nstep"])
    r = m.residual_loss(u, b).item()
    assert abs(r - float(golden["residual"])) <= 0.05 * float(golden["residual"]) + 1e-7
    # and it is a fixed point of the CUDA layer to the requested tolerance
    with torch.no_grad():
        fx = m.deqdss.f(out["result"], h0, b)
    assert float((fx - out["result"]).norm() / fx.norm()) < 1.5 * eps


def test_model_inference_matches_reference(golden):
    m = golden.model(DEV)
    b = golden.batch(DEV)
    u = m.inference(b)
    assert u.shape == (b.num_nodes, 1)
    assert rel_err(u, golden.t("u")) <= band_u(golden)


def test_broyden_generic_callable_matches_fused(golden):
    """The step API driven by a Python callable takes the same steps as the fused loop."""
    from psi_gnn_b200 import solver as S
    m = golden.model(DEV)
    b = golden.batch(DEV)
    h0 = golden.t("h0", DEV)
    op = S.LayerOperator(m.deqdss.f, h0, b)
    thr = 40
    a = S.broyden(op, h0, threshold=thr, eps=1e-30)
    c = S.broyden(lambda H: op(H), h0, threshold=thr, eps=1e-30)
    assert a["steps_run"] == c["steps_run"] == thr
    assert np.allclose(a["rel_trace"][:10], c["rel_trace"][:10], rtol=1e-3)
    assert rel_err(a["result"], c["result"]) < 1e-2


def test_picard_matches_reference(golden):
    from psi_gnn_b200 import solver as S
    m = golden.model(DEV)
    b = golden.batch(DEV)
    h0 = golden.t("h0", DEV)
    out = S.forward_iteration(S.LayerOperator(m.deqdss.f, h0, b), h0, eps=1e-4, threshold=60)
    assert out["nstep"] == int(golden["picard_nstep"])
    ref = golden["picard_rel_trace"]
    got = np.asarray(out["rel_trace"])
    assert got.shape == ref.shape
    assert np.all(np.abs(got - ref) <= 1e-3 * ref + 1e-9)
    assert rel_err(out["result"], golden.t("picard_result")) <= 1e-4


def test_anderson_matches_reference(golden):
    from psi_gnn_b200 import solver as S
    m = golden.model(DEV)
    b = golden.batch(DEV)
    h0 = golden.t("h0", DEV)
    out = S.anderson(S.LayerOperator(m.deqdss.f, h0, b), h0, m=2, threshold=60, eps=1e-4)
    ref = golden["anderson_rel_trace"]
    got = np.asarray(out["rel_trace"])
    assert got.shape == ref.shape
    # Anderson(2) is far less chaotic than Broyden: the first 8 residuals follow the reference to 1e-3; the result is an eps = 1e-4
    # accurate fixed point like the reference's (two such points differ by ~ eps / (1 − ρ))
    k = 8
    assert np.all(np.abs(got[:k] - ref[:k]) <= 1e-3 * ref[:k] + 1e-9)
    assert out["lowest"] < 1e-4 or float(golden["anderson_lowest"]) >= 1e-4
    assert rel_err(out["result"], golden.t("anderson_result")) <= 3e-3
    assert abs(out["nstep"] - int(golden["anderson_nstep"])) <= max(3, 0.15 * int(golden["anderson_nstep"]))


def _train_step(m, b, v, monkeypatch):
    from psi_gnn_b200 import model as PM
    monkeypatch.setattr(PM.torch, "randn", lambda *a, **k: v.clone())
    m.train()
    m.zero_grad()
    u, loss_dic = m(b)
    loss = loss_dic["residual_loss"].mean() + 1.0 * loss_dic["jacobian_loss"].mean() + loss_dic["encoder_loss"].mean() + \
        loss_dic["autoencoder_loss"].mean()
    loss.backward()
    monkeypatch.undo()
    gs = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).double().cpu() for _, p in m.named_parameters()])
    return u, loss_dic, gs


def _ref_grads(golden, m):
    return torch.cat([golden.t("train_grad." + k).reshape(-1).double() for k, _ in m.named_parameters()])


def test_training_step_teacher_forced(golden, monkeypatch):
    """ModelDEQDSS.forward + loss.backward() with the forward fixed point pinned to the reference's H* (so that the chaotic scatter of
    the forward trajectory is out of the picture): losses, the implicit-adjoint solve and all parameter gradients must match the
    reference's training step (same Hutchinson probe v)."""
    from psi_gnn_b200 import solver as S
    m = golden.model(DEV)
    b = golden.batch(DEV)
    hstar = golden.t("train_hstar", DEV)
    calls = []

    def solver(f, x0, threshold, eps):
        if not calls:                                  # forward solve → the reference's fixed point
            calls.append("fw")
            return {"result": hstar.clone(), "lowest": float(golden["fw_lowest"]), "nstep": int(golden["fw_nstep"]), "steps_run": 0,
                    "f_evals": 0, "launches": 0}
        calls.append("bw")
        return S.broyden(f, x0, threshold=threshold, eps=eps)

    m.deqdss.config_deq["solver"] = solver
    u, loss_dic, gs = _train_step(m, b, golden.t("train_v", DEV), monkeypatch)
    assert calls == ["fw", "bw"]
    assert rel_err(u.detach(), golden.t("train_u")) <= 1e-5
    for k in ("residual_loss", "jacobian_loss", "encoder_loss", "autoencoder_loss", "mse_loss", "mse_dirichlet"):
        ref = float(golden["train_loss." + k])
        assert abs(loss_dic[k].item() - ref) <= 1e-4 * abs(ref) + 1e-9, k
    bw = m.deqdss.last_backward
    tol = max(1e-4, 400.0 * max(bw["lowest"], float(golden["train_bw_lowest"])))
    assert rel_err(bw["result"], golden.t("train_bw_result")) <= tol
    rs = _ref_grads(golden, m)
    cos = float((gs @ rs) / (gs.norm() * rs.norm()))
    assert cos > 0.9999, cos
    assert abs(float(gs.norm() / rs.norm()) - 1) < 5e-3
    assert float((gs - rs).norm() / rs.norm()) < 2e-2


def test_training_step_free_running(golden, monkeypatch):
    """The same step with the native forward solve: the fixed point now carries the O(eps/(1−ρ)) scatter every implementation has
    (SURVEY §7.3-1), so only bands are asserted here; the tight checks are the teacher-forced tests."""
    m = golden.model(DEV)
    b = golden.batch(DEV)
    u, loss_dic, gs = _train_step(m, b, golden.t("train_v", DEV), monkeypatch)
    assert rel_err(u.detach(), golden.t("train_u")) <= band_u(golden)
    for k in ("residual_loss", "jacobian_loss", "encoder_loss", "autoencoder_loss"):
        ref = float(golden["train_loss." + k])
        assert abs(loss_dic[k].item() - ref) <= 0.25 * abs(ref) + 1e-7, k
    assert m.deqdss.last_forward["lowest"] < float(golden["cfg.fw_tol"])
    bw = m.deqdss.last_backward
    assert bw is not None and bw["lowest"] < 50 * float(golden["cfg.bw_tol"]) + float(golden["train_bw_lowest"])
    rs = _ref_grads(golden, m)
    assert torch.isfinite(gs).all()
    cos = float((gs @ rs) / (gs.norm() * rs.norm()))
    assert cos > 0.98, cos


def test_backward_solve_teacher_forced(golden):
    """The implicit-adjoint solve y = Jᵀy + grad at the reference's own H* and cotangent (both from the golden training step):
    same linear system, solved to the same tolerance by the fused VJP + Broyden kernels ⇒ same adjoint."""
    from psi_gnn_b200 import solver as S
    m = golden.model(DEV)
    b = golden.batch(DEV)
    grad = golden.t("train_bw_grad", DEV)
    op = S.VjpOperator(m.deqdss.f, golden.t("train_hstar", DEV), b, grad)
    out = S.broyden(op, torch.zeros_like(grad), threshold=int(golden["cfg.bw_thres"]), eps=float(golden["cfg.bw_tol"]))
    ref_low = float(golden["train_bw_lowest"])
    # at the 500-step cap the best residual scatters (sampled over 33 rescalings of grad on mixed_ckpt: 6e-8 … 1.3e-7 with one 2e-6
    # outlier, reference 9e-8): the band is wide, the accuracy check below scales with the residual actually reached
    assert out["lowest"] <= max(50 * ref_low, 2 * float(golden["cfg.bw_tol"]))
    # error of each solution ≈ its relative residual / (1 − ρ)
    tol = max(1e-4, 400.0 * max(out["lowest"], ref_low))
    assert rel_err(out["result"], golden.t("train_bw_result")) <= tol, (out["lowest"], ref_low)
    # and it satisfies the fixed-point equation on the CUDA operator
    y = out["result"]
    r = op(y) - y
    assert float(r.norm() / (op(y).norm() + 1e-9)) <= 3 * max(out["lowest"], float(golden["cfg.bw_tol"]))


# ---- edge cases ------------------------------------------------------------------------------------------------
def _tiny_graph(device, n, edges, mixed=False):
    from psi_gnn_b200.synthetic import GraphData
    ei = torch.tensor(edges, dtype=torch.long, device=device).t().contiguous() if edges else torch.zeros(2, 0, dtype=torch.long, device=device)
    nnz = ei.shape[1]
    gen = torch.Generator().manual_seed(7)
    b = GraphData(x=torch.zeros(n, 1, device=device), edge_index=ei,
                  edge_attr=torch.randn(nnz, 3, generator=gen).to(device), a_ij=torch.randn(nnz, 1, generator=gen).to(device),
                  y=torch.randn(n, 1, generator=gen).to(device), sol=torch.zeros(n, 1, device=device),
                  prb_data=torch.randn(n, 2, generator=gen).to(device), tags=torch.zeros(n, 1, device=device))
    b.num_nodes = n
    return b


def _cpu_f(g, h, h0, b):
    """independent fp64 CPU evaluation of the same layer by the oracle (the checker)"""
    from oracle import psignn_oracle as O
    P64 = {k: v.double() for k, v in g.params().items()}
    return O.f_dirichlet(P64, h.double().cpu(), h0.double().cpu(), b.to("cpu").double())


def test_empty_graph():
    g = Golden("dirichlet_seed0")
    m = g.model(DEV)
    b = _tiny_graph(DEV, 0, [])
    h = torch.zeros(0, 10, device=DEV)
    with torch.no_grad():
        out = m.deqdss.f(h, h, b)
    assert out.shape == (0, 10)
    res = m.deqdss.inference(h, b)
    assert res["result"].shape == (0, 10)


def test_isolated_and_ragged_nodes():
    """nodes without edges, a hub with degree 40 (ragged slices), self loops only, duplicate edges."""
    g = Golden("dirichlet_seed0")
    m = g.model(DEV)
    n = 70
    edges = [(i, i) for i in range(n)]
    edges += [(0, j) for j in range(1, 41)] + [(j, 0) for j in range(1, 41)]        # hub
    edges += [(50, 51), (51, 50), (50, 51)]                                         # duplicate edge
    edges += [(64, 33), (33, 64)]                                                   # crosses a 32-node slice boundary
    b = _tiny_graph(DEV, n, edges)
    b.tags[5] = 1.0
    b.tags[64] = 1.0
    gen = torch.Generator().manual_seed(3)
    h = torch.randn(n, 10, generator=gen).to(DEV)
    h0 = torch.randn(n, 10, generator=gen).to(DEV)
    with torch.no_grad():
        out = m.deqdss.f(h, h0, b)
    ref = _cpu_f(g, h, h0, b)
    assert rel_err(out, ref) <= TOL


def test_cpu_tensors_fail_loudly():
    g = Golden("dirichlet_seed0")
    m = g.model("cpu")
    b = g.batch("cpu")
    with pytest.raises(RuntimeError):
        m.inference(b)
    with pytest.raises(RuntimeError):
        m.deqdss.f(g.t("h0"), g.t("h0"), b)


# ---- size-independent properties at full size ----------------------------------------------------------------------
def test_large_mesh_properties():
    """C3-sized batch (256 × ~500 nodes): determinism, Dirichlet clamp, LayerNorm statistics of the free rows, and
    ⟨Jᵀy, w⟩ = ⟨y, Jw⟩ between the VJP kernels and the autograd path."""
    from psi_gnn_b200 import synthetic
    from psi_gnn_b200.solver import VjpOperator
    g = Golden("dirichlet_ckpt")
    m = g.model(DEV)
    one = synthetic.make_batch(8, seed0=100)
    b = synthetic.collate([one] * 32).to(DEV)           # 256 graphs
    with torch.no_grad():
        h0 = m._encode_native(b.x)
        f1 = m.deqdss.f(h0, h0, b)
        f1b = m.deqdss.f(h0, h0, b)
    assert torch.equal(f1, f1b)
    d = b.tags.reshape(-1) == 1
    assert torch.equal(f1[d], h0[d])
    free = f1[~d]
    gamma, beta = m.deqdss.f.laynorm.weight, m.deqdss.f.laynorm.bias
    z = (free - beta) / gamma
    assert float(z.mean(1).abs().max()) < 1e-4
    # replicated graphs give replicated outputs (no cross-talk between graphs of a batch)
    n1 = one.num_nodes
    assert torch.equal(f1[:n1], f1[n1:2 * n1])
    y, w = torch.randn_like(f1), torch.randn_like(f1)
    op = VjpOperator(m.deqdss.f, f1, b, torch.zeros_like(y))
    lhs = (op(y).double() * w.double()).sum()
    Hr = f1.clone().requires_grad_()
    with torch.enable_grad():
        out = m.deqdss.f(Hr, h0, b)
    rhs = (torch.autograd.grad(out, Hr, y)[0].double() * w.double()).sum()
    assert abs(lhs - rhs) <= 1e-4 * abs(rhs) + 1e-5


def test_c5_size_mesh_layer_vs_oracle_and_properties():
    """BASELINE configs[4] at FULL size — one 1 005 496-node mesh (7 M matrix entries), reordered as bench.py does:
      * one layer application through the C ABI vs the CPU oracle (fp32) on the same inputs: rel L2 ≤ 1e-5;
      * bit-equal re-run, Dirichlet rows bit-equal to h0, LayerNorm statistics of the free rows;
      * one VJP application vs the oracle's autograd VJP (≤ 1e-5) and bit-equal on a second run;
      * a 25-step Broyden solve: bit-equal on a second run, residual decreasing, rel trace of the first 5 steps vs the oracle's solve."""
    from oracle import psignn_oracle as O
    from psi_gnn_b200 import partition, solver as S, synthetic
    g = Golden("dirichlet_ckpt")
    m = g.model(DEV)
    host = partition.reorder_mesh(synthetic.make_large_mesh(1_000_000, seed=0))
    assert host.num_nodes > 1_000_000
    b = host.to(DEV)
    P = g.params()
    with torch.no_grad():
        h0 = m._encode_native(b.x)
        f1 = m.deqdss.f(h0, h0, b)
        f1b = m.deqdss.f(h0, h0, b)
        f2 = m.deqdss.f(f1, h0, b)
    assert torch.equal(f1, f1b)
    h0_ref = O.encoder(P, host.x)
    assert rel_err(h0, h0_ref) <= TOL
    Hc = f1.cpu().requires_grad_()
    ref2 = O.f_dirichlet(P, Hc, h0.cpu(), host)                       # second application, on the CUDA path's own f1
    assert rel_err(f2, ref2.detach()) <= TOL, rel_err(f2, ref2.detach())
    d = b.tags.reshape(-1) == 1
    assert torch.equal(f1[d], h0[d]) and int(d.sum()) > 1000
    z = (f1[~d] - m.deqdss.f.laynorm.bias) / m.deqdss.f.laynorm.weight
    assert float(z.mean(1).abs().max()) < 1e-4
    # VJP at the same point.  Among the 1.2e8 ReLUs of a 1 M-node mesh a few pre-activations sit within fp32 rounding of zero, so ANY
    # fp32 evaluation flips a few masks against the fp64 one (a derivative discontinuity): the oracle's own fp32 VJP is 9e-5 away from
    # its fp64 VJP here.  Band: our distance to the fp64 VJP ≤ 3 × the fp32 oracle's, and the error sits in a handful of rows.
    y = torch.randn(f1.shape, generator=torch.Generator().manual_seed(7))
    v32 = torch.autograd.grad(ref2, Hc, y)[0]
    P64 = {k_: v_.double() for k_, v_ in P.items()}
    H64 = f1.cpu().double().requires_grad_()
    v64 = torch.autograd.grad(O.f_dirichlet(P64, H64, h0.cpu().double(), host.double()), H64, y.double())[0]
    op = S.VjpOperator(m.deqdss.f, f1, b, torch.zeros_like(f1))
    v1, v2 = op(y.to(DEV)), op(y.to(DEV))
    assert torch.equal(v1, v2)
    e_ours, e_ref = rel_err(v1, v64), rel_err(v32, v64)
    row_err = (v1.cpu().double() - v64).norm(dim=1) / (v64.norm(dim=1) + 1e-12)
    bad_rows = int((row_err > 1e-4).sum())
    print("C5-size VJP vs fp64 oracle: ours %.2e, fp32 oracle %.2e; rows off by > 1e-4: %d of %d" % (e_ours, e_ref, bad_rows, v64.shape[0]))
    assert e_ours <= 3.0 * e_ref + TOL, (e_ours, e_ref)
    assert bad_rows <= 1e-4 * v64.shape[0], bad_rows
    assert float(torch.median(row_err)) <= TOL
    lop = S.LayerOperator(m.deqdss.f, h0, b)
    r1 = S.broyden(lop, h0, threshold=25, eps=1e-30)
    r2 = S.broyden(lop, h0, threshold=25, eps=1e-30)
    assert r1["rel_trace"] == r2["rel_trace"] and torch.equal(r1["result"], r2["result"])
    assert r1["rel_trace"][-1] < 0.2 * r1["rel_trace"][0]
    # the oracle's solve (reference algorithm, fp32, CPU).  Its own rel values come from fp32 norms over 1e7 elements, which torch's
    # CPU reduction gets right only to ≈ 1e-3 (0.0025252 vs 0.0025284 with fp64 accumulation, thread-count dependent), so the residuals
    # of ITS iterates are re-measured with fp64 norms; the CUDA path (fp64 sums of per-block fp32 partials) must match those.
    rec = []

    def f_oracle(H):
        Y = O.f_dirichlet(P, H, h0_ref, host)
        rec.append(float((Y - H).double().norm() / Y.double().norm()))
        return Y

    ref = O.broyden(f_oracle, h0_ref, 5, 1e-30)
    want = np.asarray(rec[1:6])                      # rel_trace[j] belongs to the (j+1)-th evaluation
    got = np.asarray(r1["rel_trace"][:len(want)])
    assert np.all(np.abs(np.asarray(ref["rel_trace"][:len(want)]) - want) <= 5e-3 * want)
    assert np.all(np.abs(got - want) <= 1e-3 * want), (got, want)


# ---- DSS / DSGPS baselines on the shared layer kernel (config 2) ------------------------------------------------------
def _baseline(name):
    g = Golden(name)
    z = g.z
    cfg = dict(latent_dim=10, k=int(z["cfg.k"]), alpha=float(z["cfg.alpha"]), gamma=float(z["cfg.gamma"]) if "cfg.gamma" in z.files else 0.9)
    if name.startswith("dss"):
        from psi_gnn_b200.dirichlet.dss import model as M
        m = M.DeepStatisticalSolver(cfg)
    elif "mixed" in name:
        from psi_gnn_b200.mixed.dsgps import model as M
        m = M.ModelDSGPS(cfg)
    else:
        from psi_gnn_b200.dirichlet.dsgps import model as M
        m = M.ModelDSGPS(cfg)
    m.load_state_dict(g.params())
    from psi_gnn_b200.synthetic import GraphData
    b = GraphData()
    for k in z.files:
        if k.startswith("batch.") and k != "batch.num_nodes":
            setattr(b, k[6:], g.t(k, DEV))
    b.num_nodes = int(z["batch.num_nodes"])
    return g, m.to(DEV), b


def test_dss_inference_matches_reference():
    g, m, b = _baseline("dss_ckpt")
    u = m.inference(b)
    assert rel_err(u, g.t("u")) <= TOL


def test_dss_single_layer_matches_reference():
    from psi_gnn_b200 import _native as N, weights as W
    from psi_gnn_b200.graph import graph_of
    g, m, b = _baseline("dss_ckpt")
    k = int(g["layer_index"])
    blob = W.pack_dss(g.params(DEV), k, float(g["cfg.alpha"]), DEV)
    W.upload(blob, W.next_serial())
    out = graph_of(b, N.KIND_DSS).layer_forward(N.KIND_DSS, g.t("layer_in", DEV), None)
    assert rel_err(out, g.t("layer_out")) <= TOL
    # the increment itself (α = 1e-3 hides errors in H + α·Ψ): compare Ψ = (out − in)/α to 1e-4
    inc = (out - g.t("layer_in", DEV)).double().cpu()
    ref = (g.t("layer_out") - g.t("layer_in")).double()
    assert float((inc - ref).norm() / ref.norm()) <= 1e-3


def test_dsgps_inference_matches_reference():
    g, m, b = _baseline("dsgps_ckpt")
    u = m.inference(b)
    assert rel_err(u, g.t("u")) <= 5 * TOL          # 30 recurrent steps


def test_dsgps_single_layer_matches_reference():
    from psi_gnn_b200 import _native as N, weights as W
    from psi_gnn_b200.graph import graph_of
    g, m, b = _baseline("dsgps_ckpt")
    W.upload(W.pack_dsgps(g.params(DEV), DEV), W.next_serial())
    out = graph_of(b, N.KIND_DSGPS).layer_forward(N.KIND_DSGPS, g.t("layer_in", DEV), g.t("layer_h0", DEV))
    assert rel_err(out, g.t("layer_out")) <= TOL


def test_dsgps_mixed_inference_and_layer_match_reference():
    """mixed/dsgps (reference mixed/dsgps/model.py:76-97): native layer kind 4 — Neumann rows overwritten by update_neumann, Dirichlet
    rows clamped — one layer and the 30-step unrolled inference against the unmodified reference"""
    from psi_gnn_b200 import _native as N, weights as W
    from psi_gnn_b200.graph import graph_of
    g, m, b = _baseline("dsgps_mixed_ckpt")
    u = m.inference(b)
    assert rel_err(u, g.t("u")) <= 5 * TOL          # 30 recurrent steps
    W.upload(W.pack_dsgps(g.params(DEV), DEV), W.next_serial())
    out = graph_of(b, N.KIND_DSGPS_MIXED).layer_forward(N.KIND_DSGPS_MIXED, g.t("layer_in", DEV), g.t("layer_h0", DEV))
    assert rel_err(out, g.t("layer_out")) <= TOL
    t = b.tags
    assert torch.equal(out[t[:, 1] == 1], g.t("layer_h0", DEV)[t[:, 1] == 1])       # Dirichlet rows are copied


def _baseline_step_oracle(name, P, hh, h0, bc, alpha, k=3):
    """one unrolled step of a baseline in the oracle's torch form (fp64 when its inputs are): DSS layer k, DSGPS step, mixed DSGPS step
    (reference dirichlet/dss/model.py:83-91, dirichlet/dsgps/model.py:143-163, mixed/dsgps/model.py:76-97)"""
    from oracle import psignn_oracle as O
    if name.startswith("dss"):
        return O.dss_layer(P, k, hh, bc, alpha)
    if "mixed" not in name:
        return O.dsgps_layer(P, hh, h0, bc)
    ei, attr = O.offdiag(bc.edge_index, bc.edge_attr)
    to, fr, ne = (O.phi(P, p_, hh, ei, attr, t_) for p_, t_ in (("phi_to", True), ("phi_from", False), ("phi_neumann", False)))
    c = torch.cat([hh, to, fr, bc.prb_data], 1)
    zg, rg = torch.sigmoid(O._lin(P, "z_k.mlp.0", c)), torch.sigmoid(O._lin(P, "r_k.mlp.0", c))
    corr = torch.tanh(O._lin(P, "correction.mlp.0", torch.cat([rg * hh, to, fr, bc.prb_data], 1)))
    upd = O.mlp2(P, "update_neumann.mlp", torch.cat([hh, ne, bc.prb_data, bc.unit_normal_vector], 1))
    out = torch.where((bc.tags[:, 2] == 1)[:, None], upd, hh + zg * corr)
    return torch.where((bc.tags[:, 1] == 1)[:, None], h0, out)


@pytest.mark.parametrize("name", ["dss_ckpt", "dsgps_ckpt", "dsgps_mixed_ckpt"])
@pytest.mark.parametrize("seed,n", [(0, 33), (2, 257), (3, 1000)])
def test_random_ragged_graphs_baseline_backward(name, seed, n):
    """psi_layer_forward / psi_layer_backward of the baseline kinds on randomised ragged multigraphs (isolated nodes, a hub of degree
    n/3, self loops, duplicate and asymmetric edges, every boundary class) against the oracle's fp64 step and autograd through it"""
    from psi_gnn_b200 import _native as N, weights as W
    from psi_gnn_b200.graph import graph_of
    g, m, _ = _baseline(name)
    mixed = "mixed" in name
    b, h, h0, y = _random_graph(seed, n, mixed, "cpu")
    if name.startswith("dss"):
        gen = torch.Generator().manual_seed(100 + seed)
        b.a_ij_norm = torch.randn(b.edge_index.shape[1], 1, generator=gen)
        b.b_prime_norm = torch.randn(n, 3, generator=gen)
        b.b_prime = torch.randn(n, 3, generator=gen)
        kind, k = N.KIND_DSS, 3
        block, unpack = m._layer_block(k, DEV), (lambda flat: W.unpack_dss_grads(flat, k))
    else:
        kind = N.KIND_DSGPS_MIXED if mixed else N.KIND_DSGPS
        block, unpack = m._layer_block(0, DEV), (lambda flat: W.unpack_dsgps_grads(flat, mixed))
    P = {k_: v.double().requires_grad_() for k_, v in g.params().items()}
    hh = h.double().requires_grad_()
    out = _baseline_step_oracle(name, P, hh, h0.double(), b.double(), m.config["alpha"])
    bd = b.to(DEV)
    gr = graph_of(bd, kind)
    W.upload(*block)
    f_native = gr.layer_forward(kind, h.to(DEV), None if kind == N.KIND_DSS else h0.to(DEV))
    assert rel_err(f_native, out.detach()) <= TOL
    hbar, flat = gr.layer_backward(kind, h.to(DEV), y.to(DEV))
    grads = unpack(flat)
    names = list(grads)
    ref = torch.autograd.grad(out, [hh] + [P[k_] for k_ in names], y.double(), allow_unused=True)
    assert rel_err(hbar, ref[0]) <= TOL, rel_err(hbar, ref[0])
    got = torch.cat([grads[k_].reshape(-1).double().cpu() for k_ in names])
    want = torch.cat([(r if r is not None else torch.zeros_like(P[k_])).reshape(-1) for k_, r in zip(names, ref[1:])])
    assert float((got - want).norm() / want.norm()) <= TOL, float((got - want).norm() / want.norm())


@pytest.mark.parametrize("name", ["dss_ckpt", "dsgps_ckpt", "dsgps_mixed_ckpt"])
def test_baseline_layer_backward_matches_autograd(name):
    """psi_layer_backward (one unrolled DSS / DSGPS / mixed DSGPS step: h̄ = Jᵀȳ and every parameter gradient) through the C ABI
    against autograd through the oracle's fp64 form of the same step (reference dirichlet/dss/model.py:83-91, dirichlet/dsgps/
    model.py:143-163, mixed/dsgps/model.py:76-97) at a random point; bit-equal on a second call (no atomics)"""
    from oracle import psignn_oracle as O
    from psi_gnn_b200 import _native as N, weights as W
    from psi_gnn_b200.graph import graph_of
    g, m, b = _baseline(name)
    cfg = m.config
    n = b.num_nodes
    gen = torch.Generator().manual_seed(3)
    h = torch.randn(n, 10, generator=gen) * 0.5
    y = torch.randn(n, 10, generator=gen)
    h0 = torch.randn(n, 10, generator=gen)
    P = {k: v.double().requires_grad_() for k, v in g.params().items()}
    bc = g.batch()
    for a in ("edge_attr", "prb_data", "unit_normal_vector", "a_ij_norm", "b_prime_norm"):
        if getattr(bc, a, None) is not None:
            setattr(bc, a, getattr(bc, a).double())
    hh = h.double().requires_grad_()
    if name.startswith("dss"):
        kind, k = N.KIND_DSS, 3
        out = O.dss_layer(P, k, hh, bc, cfg["alpha"])
        W.upload(*m._layer_block(k, DEV))
        unpack = lambda flat: W.unpack_dss_grads(flat, k)
    else:
        mixed = "mixed" in name
        kind = N.KIND_DSGPS_MIXED if mixed else N.KIND_DSGPS
        if mixed:
            ei, attr = O.offdiag(bc.edge_index, bc.edge_attr)
            to, fr, ne = (O.phi(P, p_, hh, ei, attr, t_) for p_, t_ in (("phi_to", True), ("phi_from", False), ("phi_neumann", False)))
            c = torch.cat([hh, to, fr, bc.prb_data], 1)
            zg, rg = torch.sigmoid(O._lin(P, "z_k.mlp.0", c)), torch.sigmoid(O._lin(P, "r_k.mlp.0", c))
            corr = torch.tanh(O._lin(P, "correction.mlp.0", torch.cat([rg * hh, to, fr, bc.prb_data], 1)))
            upd = O.mlp2(P, "update_neumann.mlp", torch.cat([hh, ne, bc.prb_data, bc.unit_normal_vector], 1))
            out = torch.where((bc.tags[:, 2] == 1)[:, None], upd, hh + zg * corr)
            out = torch.where((bc.tags[:, 1] == 1)[:, None], h0.double(), out)
        else:
            out = O.dsgps_layer(P, hh, h0.double(), bc)
        W.upload(*m._layer_block(0, DEV))
        unpack = lambda flat: W.unpack_dsgps_grads(flat, mixed)
    gr = graph_of(b, kind)
    # the native forward of the same step first (the masks of the backward are the forward's)
    f_native = gr.layer_forward(kind, h.to(DEV), None if kind == N.KIND_DSS else h0.to(DEV))
    assert rel_err(f_native, out.detach().float()) <= TOL
    hbar, flat = gr.layer_backward(kind, h.to(DEV), y.to(DEV))
    hbar2, flat2 = gr.layer_backward(kind, h.to(DEV), y.to(DEV))
    assert torch.equal(hbar, hbar2) and torch.equal(flat, flat2)
    grads = unpack(flat)
    names = list(grads)
    ref = torch.autograd.grad(out, [hh] + [P[k_] for k_ in names], y.double(), allow_unused=True)
    assert rel_err(hbar, ref[0].float()) <= TOL, rel_err(hbar, ref[0].float())
    got = torch.cat([grads[k_].reshape(-1).double().cpu() for k_ in names])
    want = torch.cat([(r if r is not None else torch.zeros_like(P[k_])).reshape(-1) for k_, r in zip(names, ref[1:])])
    assert float((got - want).norm() / want.norm()) <= TOL, float((got - want).norm() / want.norm())
    for k_, r in zip(names, ref[1:]):
        r = torch.zeros_like(P[k_]) if r is None else r
        assert float((grads[k_].double().cpu() - r).norm()) <= 1e-4 * float(r.norm()) + 1e-6 * float(want.norm()), k_


def test_dss_flux_residual_native_matches_torch():
    """DSS flux-form residual (reference dirichlet/dss/model.py:129-148): the native edge sums (psi_flux, forward and adjoint) against
    the reference-signature torch form — value and gradient w.r.t. U; bit-equal on a second call (no atomics)"""
    g, m, b = _baseline("dss_ckpt")
    gen = torch.Generator().manual_seed(5)
    U = torch.randn(b.num_nodes, 1, generator=gen).to(DEV).requires_grad_()
    a = m._residual_native(U, b)
    (ga,) = torch.autograd.grad(a, U)
    U2 = U.detach().clone().requires_grad_()
    r = m.residual_loss(U2.double(), b.edge_index, b.a_ij.double(), b.b_prime.double())
    (gr,) = torch.autograd.grad(r, U2)
    assert abs(a.item() - r.item()) <= 1e-6 * abs(r.item())
    assert rel_err(ga, gr) <= TOL
    a2 = m._residual_native(U, b)
    (ga2,) = torch.autograd.grad(a2, U)
    assert torch.equal(a, a2) and torch.equal(ga, ga2)


class _TorchBackwardLayer(torch.autograd.Function):
    """test double of baselines._UnrolledLayer: the same native forward, but the backward is torch autograd through the differentiable
    torch form of the step, recomputed at the saved input — the reference's backward evaluated at the native path's own states"""

    @staticmethod
    def forward(ctx, owner, batch, step, block, dmask, h, h0, *params):
        from psi_gnn_b200 import weights as W
        from psi_gnn_b200.graph import graph_of
        W.upload(*block)
        ctx.owner, ctx.batch, ctx.step, ctx.has_h0 = owner, batch, step, h0 is not None
        ctx.save_for_backward(h, h0 if h0 is not None else h)
        return graph_of(batch, owner._layer_kind).layer_forward(owner._layer_kind, h.detach(), h0.detach() if h0 is not None else None)

    @staticmethod
    def backward(ctx, ybar):
        h, h0 = ctx.saved_tensors
        owner = ctx.owner
        P = dict(owner.named_parameters())
        with torch.enable_grad():
            h_, h0_ = h.detach().requires_grad_(), h0.detach().requires_grad_()
            out = owner._step_torch(ctx.step, h_, h0_ if ctx.has_h0 else None, ctx.batch)
            gr = torch.autograd.grad(out, [h_] + ([h0_] if ctx.has_h0 else []) + [P[n] for n in owner._layer_names(ctx.step)], ybar, allow_unused=True)
        return (None,) * 5 + ((gr[0], gr[1]) + tuple(gr[2:]) if ctx.has_h0 else (gr[0], None) + tuple(gr[1:]))


# band of the whole-step gradient against the reference's own fp32 run.  An unrolled baseline takes ~2.4 M ReLU decisions per training
# step; a pre-activation within fp32 rounding of zero is decided differently by ANY two fp32 evaluations (and by the fp64 one), and one
# flipped unit moves the gradient by ~1e-3 (a derivative discontinuity, not an error: profiles/r02_baseline_backward.md — on dsgps_ckpt the
# cotangents of steps 21 → 20 jump from 2.5e-4 to 2.2e-3 off the fp64 truth with 99 % of that in the rows of ONE edge, while the states
# agree to 5e-7 and the torch-autograd backward evaluated at the same native states reproduces the native gradient to 3e-6).  The strict
# statement is therefore the second one: native backward vs torch autograd at identical states.
_GOLDEN_GRAD_BAND = {"dss_ckpt": 1e-4, "dsgps_ckpt": 2e-3, "dsgps_mixed_ckpt": 1e-4}       # measured: 1.9e-6, 9.4e-4 (one flip), 9.5e-7


@pytest.mark.parametrize("name", ["dss_ckpt", "dsgps_ckpt", "dsgps_mixed_ckpt"])
def test_baseline_training_step_and_checkpoint_roundtrip(name, tmp_path, monkeypatch):
    """unrolled training forward + backward of the baselines (reference dirichlet/dss/model.py:59-104, */dsgps/model.py:48-131) on the
    native layer kernels (forward: psi_layer_forward, backward: psi_layer_backward): total loss, last state and every parameter gradient
    against the unmodified reference; the native backward against torch autograd through the torch form of every step at the SAME
    states (≤ 2e-5 over all parameters), bit-equal on a second run; then a checkpoint in the reference's layout (training_class.py:297-307:
    epoch, hyperparameters, state_dict, …) survives torch.save / torch.load / load_state_dict and the reloaded model reproduces the
    native inference bit for bit"""
    from psi_gnn_b200 import baselines as B
    g, m, b = _baseline(name)
    m.train()

    def step():
        m.zero_grad()
        U, ld = m(b)
        ld["train_loss"].backward()
        return U, ld, torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).double().cpu() for _, p in m.named_parameters()])

    U, ld, gs = step()
    _, _, gs2 = step()
    assert torch.equal(gs[: gs.numel()], gs2) or float((gs - gs2).norm() / gs.norm()) < 1e-6      # layers are deterministic; the torch loss block need not be
    k = str(m.config["k"])
    ref = float(g["train_loss"])
    assert abs(ld["train_loss"].item() - ref) <= 2e-5 * abs(ref)
    assert rel_err(U[k].detach(), g.t("train_u_last")) <= 5 * TOL
    assert abs(ld["residual_loss"][k].item() - float(g["train_res_last"])) <= 1e-4 * abs(float(g["train_res_last"]))
    rs = torch.cat([g.t("train_grad." + n).reshape(-1).double() for n, _ in m.named_parameters()])
    err_golden = float((gs - rs).norm() / rs.norm())
    monkeypatch.setattr(B, "_UnrolledLayer", _TorchBackwardLayer)
    _, _, gt = step()
    monkeypatch.undo()
    err_same_states = float((gs - gt).norm() / gt.norm())
    print("%s: gradient vs the reference's fp32 run %.2e, vs torch autograd at the native states %.2e" % (name, err_golden, err_same_states))
    assert err_same_states <= 2e-5, err_same_states
    assert err_golden <= _GOLDEN_GRAD_BAND[name], err_golden
    # checkpoint round-trip
    path = tmp_path / "best_model.pt"
    torch.save({"epoch": 3, "hyperparameters": dict(m.config), "state_dict": m.state_dict(), "hist_train": {}, "hist_val": {}, "training_time": 1.0}, path)
    ck = torch.load(path, map_location=DEV, weights_only=False)
    m2 = type(m)(ck["hyperparameters"]).to(DEV)
    m2.load_state_dict(ck["state_dict"])
    assert set(ck["state_dict"].keys()) == set(g.params().keys())
    m.eval(); m2.eval()
    assert torch.equal(m2.inference(b), m.inference(b))


def test_evaluation_form_models():
    """ModelPSIGNN / ModelPSIGNNIterative (reference tests/model_psignn.py:28-214): the evaluation-form wrappers around the same solve"""
    from psi_gnn_b200.dirichlet.psignn import model as M
    from psi_gnn_b200.dirichlet.psignn.utilities import solver as S
    g = Golden("dirichlet_ckpt")
    cfg = g.cfg()
    cfg["solver"] = S.broyden
    b = g.batch(DEV)
    m = M.ModelPSIGNN(cfg)
    m.load_state_dict(g.params())
    m = m.to(DEV)
    u, ld = m(b)
    assert rel_err(u, g.t("u")) <= band_u(g)
    assert ld["nsteps"] == m.deqdss.last_forward["nstep"] and abs(ld["nsteps"] - int(g["fw_nstep"])) <= 0.25 * int(g["fw_nstep"])
    assert abs(ld["residual_loss"].item() - float(g["residual"])) <= 0.05 * float(g["residual"])
    assert set(ld) == {"residual_loss", "encoder_loss", "autoencoder_loss", "mse_loss", "mse_dirichlet_loss", "nsteps"}
    mi = M.ModelPSIGNNIterative(cfg)
    mi.load_state_dict(g.params())
    out = mi.to(DEV)(b)
    assert len(out["sol_dic"]) == mi.deqdss.last_forward["steps_run"] + 2 and out["nstep"] == mi.deqdss.last_forward["nstep"]


def test_training_loop_on_the_dropin_modules(tmp_path):
    """the reference's TrainModel step recipe (dirichlet/psignn/training_class.py:147-166) driven through the PyG stand-ins
    (DataListLoader → DataParallel.collate → ModelDEQDSS): two epochs over 3 batches of 4 meshes; the collated device batches and
    their native re-layouts are built once (cache hits on the second epoch), the parameters move, the checkpoint round-trips"""
    from psi_gnn_b200 import synthetic, training as T
    from psi_gnn_b200.dirichlet.psignn import model as M
    from psi_gnn_b200.dirichlet.psignn.utilities import solver as S
    g = Golden("dirichlet_ckpt")
    cfg = g.cfg()
    cfg["solver"] = S.broyden
    net = M.ModelDEQDSS(cfg)
    net.load_state_dict(g.params())
    model = T.DataParallel(net.to(DEV), device=DEV)
    dataset = [synthetic.make_mesh(200 + i, h=0.11) for i in range(12)]
    loader = T.DataListLoader(dataset, batch_size=4, shuffle=False)
    tm = T.TrainModel({"loader_train": loader, "loader_val": None, "model": model, "config_model": cfg, "lr_deq": 1e-4, "lr_ae": 1e-4,
                       "sched_step_deq": 0.5, "sched_step_ae": 0.5, "path_ckpt": str(tmp_path), "max_epochs": 2, "gradient_clip": 0.1,
                       "jac_weight": 1.0, "sup_weight": 0.0, "min_loss_save": 1e9})
    before = torch.cat([p.detach().reshape(-1).clone() for p in net.parameters()])
    hist_train, _ = tm.train_model()
    after = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    assert len(hist_train["loss"]) == 2 and all(np.isfinite(v) for v in hist_train["loss"])
    assert float((after - before).abs().max()) > 0
    assert model.cache_hits == 3                                   # second epoch: collation, H2D copy and SELL build all skipped
    ck = torch.load(tmp_path / "final_model.pt", map_location=DEV, weights_only=False)
    assert set(ck) >= {"epoch", "hyperparameters", "state_dict", "hist_train", "hist_val", "opt_deq", "opt_ae", "sched_deq", "sched_ae", "training_time"}
    net2 = M.ModelDEQDSS(cfg)
    net2.load_state_dict(ck["state_dict"])
    b = T.Batch.from_data_list(dataset[:2]).to(DEV)
    assert torch.equal(net2.to(DEV).inference(b), net.inference(b))


def test_stale_graph_cache_is_rebuilt():
    """the native handle snapshots tags / prb_data / a_ij / edge_attr: reassigning or modifying them in place on the same batch (new
    right-hand side on the same mesh) must rebuild it — otherwise the solves and the autograd path would see different functions"""
    from psi_gnn_b200.graph import graph_of, invalidate
    g = Golden("dirichlet_seed0")
    m = g.model(DEV)
    b = g.batch(DEV)
    h0 = g.t("h0", DEV)
    with torch.no_grad():
        a = m.deqdss.f(h0, h0, b)
        g1 = graph_of(b, 0)
        b.prb_data.mul_(2.0)                                         # in place: same storage, new version
        c = m.deqdss.f(h0, h0, b)
        assert graph_of(b, 0) is not g1 and not torch.equal(a, c)
        b.prb_data = b.prb_data / 2.0                                # reassigned
        d = m.deqdss.f(h0, h0, b)
    assert rel_err(d, a) <= 1e-6
    g2 = graph_of(b, 0)
    b.prb_data.data.mul_(3.0)                                        # .data writes carry no version: explicit invalidation
    assert graph_of(b, 0) is g2
    invalidate(b)
    assert graph_of(b, 0) is not g2


def test_edge_index_out_of_range_fails_loudly():
    from psi_gnn_b200.graph import graph_of
    g = Golden("dirichlet_seed0")
    b = g.batch(DEV)
    b.edge_index = b.edge_index.clone()
    b.edge_index[1, 5] = b.num_nodes + 3
    with pytest.raises(RuntimeError, match="out of range"):
        graph_of(b, 0)


def test_keep_trace_on_every_solver():
    """the reference's solvers always return the iterates; here keep_trace=True asks for them (iterative_inference does) on Broyden,
    Picard and Anderson alike, fused and callable paths"""
    from psi_gnn_b200 import solver as S
    g = Golden("dirichlet_seed0")
    m = g.model(DEV)
    b = g.batch(DEV)
    h0 = g.t("h0", DEV)
    op = S.LayerOperator(m.deqdss.f, h0, b)
    p = S.forward_iteration(op, h0, eps=1e-4, threshold=40, keep_trace=True)
    assert len(p["xest_trace"]) == p["nstep"] + 2 and torch.equal(p["xest_trace"][0], h0) and torch.equal(p["xest_trace"][-1], p["result"])
    pc = S.forward_iteration(lambda H: op(H), h0, eps=1e-4, threshold=40, keep_trace=True)
    assert len(pc["xest_trace"]) == len(p["xest_trace"]) and rel_err(pc["xest_trace"][3], p["xest_trace"][3]) < 1e-6
    a = S.anderson(op, h0, m=2, threshold=30, eps=1e-4, keep_trace=True)
    assert len(a["xest_trace"]) == a["steps_run"] + 1 and torch.equal(a["xest_trace"][0], h0)
    assert torch.equal(a["xest_trace"][-1], a["result"])            # the trace holds the best iterate so far (solver.py:273)
    assert S.anderson(op, h0, m=2, threshold=30, eps=1e-4)["xest_trace"] == []


@pytest.mark.timeout(600)
def test_mesh_partitioned_solve_two_gpus():
    """BASELINE config 5 at test size: one mesh split over 2 ranks at 2048-node-aligned cuts (halo exchange + all-reduced inner
    products) must retrace the single-GPU solve of the same reordered mesh step for step (identical step counts, rel-trace deviation
    < 1e-6, u within 1e-5).  Needs 2 GPUs; on a 1-GPU box the host logic is covered by tests/test_multi_gloo.py."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from conftest import ROOT
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu", "partitioned_solve.py"), "12000"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=550)
    assert p.returncode == 0 and "PARTITIONED_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


# ---- adjacent API surface --------------------------------------------------------------------------------
def test_eval_mode_spectral_radius(tmp_path):
    """DeepEquilibrium.forward under no_grad (validation branch, model.py:228-241): Jacobian estimate + 150 power iterations on the
    native VJP; the trained checkpoint is a contraction with spectral radius just below 1 (reference logs: ≈ 0.990)."""
    g = Golden("dirichlet_ckpt")
    m = g.model(DEV).eval()
    m.deqdss.path_logs = str(tmp_path)
    b = g.batch(DEV)
    with torch.no_grad():
        h0 = m._encode_native(b.x)
        new_h, jac = m.deqdss(h0, b)
    assert new_h.shape == h0.shape and float(jac) > 0
    rho = float(open(tmp_path / "spectral_radius.csv").read().strip().split()[-1])
    assert 0.9 < rho < 1.01, rho


def test_iterative_inference_trace():
    g = Golden("dirichlet_ckpt")
    m = g.model(DEV)
    b = g.batch(DEV)
    out = m.iterative_inference(b)
    steps = m.deqdss.last_forward["steps_run"]
    assert len(out["sol_dic"]) == steps + 2 == len(out["res_dic"])          # batch.x, x0, then every iterate
    assert out["nstep"] == m.deqdss.last_forward["nstep"]
    assert out["res_dic"][-1] < out["res_dic"][1]                            # the residual went down along the trace
    u = m.inference(b)
    assert rel_err(out["sol_dic"][1 + out["nstep"]].to(DEV), u) < 1e-5       # the best iterate is the one inference returns


def test_forward_iteration_generic_callable():
    from psi_gnn_b200 import solver as S
    g = Golden("dirichlet_seed0")
    m = g.model(DEV)
    b = g.batch(DEV)
    h0 = g.t("h0", DEV)
    op = S.LayerOperator(m.deqdss.f, h0, b)
    a = S.forward_iteration(op, h0, eps=1e-4, threshold=40)
    c = S.forward_iteration(lambda H: op(H), h0, eps=1e-4, threshold=40)
    assert a["nstep"] == c["nstep"]
    assert rel_err(a["result"], c["result"]) < 1e-6


# ---- randomised ragged graphs against the fp64 oracle (both families, layer and VJP) ----------------------------------------------
def _random_graph(seed, n, mixed, device):
    """directed multigraph with isolated nodes, hubs, self loops, asymmetric edges, all three boundary classes"""
    from psi_gnn_b200.synthetic import GraphData
    gen = torch.Generator().manual_seed(seed)
    m = int(n * 5)
    row = torch.randint(0, n, (m,), generator=gen)
    col = torch.randint(0, n, (m,), generator=gen)
    hub = torch.randint(0, n, (1,), generator=gen).item()
    row[: n // 3] = hub                                       # a hub with degree ~n/3
    diag = torch.arange(n)
    ei = torch.stack([torch.cat([row, diag]), torch.cat([col, diag])])
    nnz = ei.shape[1]
    cls = torch.randint(0, 3 if mixed else 2, (n,), generator=gen)
    cls[torch.randint(0, n, (n // 10,), generator=gen)] = 0
    if mixed:
        tags = torch.nn.functional.one_hot(cls, 3).float()
    else:
        tags = (cls == 1).float().reshape(-1, 1)
    b = GraphData(x=torch.randn(n, 1, generator=gen), edge_index=ei, edge_attr=torch.randn(nnz, 3, generator=gen),
                  a_ij=torch.randn(nnz, 1, generator=gen), y=torch.randn(n, 1, generator=gen), sol=torch.zeros(n, 1),
                  prb_data=torch.randn(n, 3 if mixed else 2, generator=gen), tags=tags)
    if mixed:
        b.unit_normal_vector = torch.randn(n, 2, generator=gen)
    b.num_nodes = n
    h = torch.randn(n, 10, generator=gen)
    h0 = torch.randn(n, 10, generator=gen)
    y = torch.randn(n, 10, generator=gen)
    return b, h, h0, y


@pytest.mark.parametrize("mixed", [False, True])
@pytest.mark.parametrize("seed,n", [(0, 33), (1, 97), (2, 257), (3, 1000)])
def test_random_ragged_graphs_layer_and_vjp(mixed, seed, n):
    from oracle import psignn_oracle as O
    from psi_gnn_b200.solver import VjpOperator
    g = Golden("mixed_ckpt" if mixed else "dirichlet_ckpt")
    m = g.model(DEV)
    b, h, h0, y = _random_graph(seed, n, mixed, "cpu")
    P64 = {k: v.double() for k, v in g.params().items()}
    f = O.f_mixed if mixed else O.f_dirichlet
    H = h.double().requires_grad_()
    ref = f(P64, H, h0.double(), b.double())
    ref_vjp = torch.autograd.grad(ref, H, y.double())[0]
    bd = b.to(DEV)
    with torch.no_grad():
        out = m.deqdss.f(h.to(DEV), h0.to(DEV), bd)
    assert rel_err(out, ref.detach()) <= TOL
    op = VjpOperator(m.deqdss.f, h.to(DEV), bd, torch.zeros(n, 10, device=DEV))
    assert rel_err(op(y.to(DEV)), ref_vjp) <= TOL
    # residual SpMV on the same ragged matrix
    u = torch.randn(n, 1, generator=torch.Generator().manual_seed(seed))
    r = m.residual_loss(u.to(DEV), bd)
    ref_r = O.residual_loss(u.double(), b.double())
    assert abs(r.item() - float(ref_r)) <= 1e-5 * float(ref_r)


def test_fused_per_edge_product_is_bit_identical_to_the_prepass(tmp_path):
    """sub-wave grids run the layer without the pre-pass launch (W1j·h_j recomputed per edge with the same rounding chain,
    layer.cuh walk_direct_h); PSI_NO_FUSED_PRE=1 restores the two-launch path.  Both must give the same bits: layer application
    (dirichlet, DSS, DSGPS) and a whole forward solve (trace and result)."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    script = tmp_path / "run.py"
    script.write_text(
        "import sys, os, torch\n"
        "sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'tests'))\n"
        "import test_gpu_parity as T\n"
        "from conftest import Golden\n"
        "from psi_gnn_b200 import _native as N, weights as W\n"
        "from psi_gnn_b200.graph import graph_of\n"
        "out = {}\n"
        "g = Golden('dirichlet_ckpt'); m = g.model('cuda:0'); b = g.batch('cuda:0')\n"
        "with torch.no_grad():\n"
        "    h0 = m._encode_native(b.x); out['psi'] = m.deqdss.f(h0, h0, b).cpu()\n"
        "    r = m.deqdss.inference(h0, b); out['solve'] = r['result'].cpu(); out['trace'] = torch.tensor(r['rel_trace'])\n"
        "for name, kind in (('dss_ckpt', N.KIND_DSS), ('dsgps_ckpt', N.KIND_DSGPS)):\n"
        "    g2, m2, b2 = T._baseline(name)\n"
        "    W.upload(*m2._layer_block(3, 'cuda:0'))\n"
        "    gen = torch.Generator().manual_seed(1); h = torch.randn(b2.num_nodes, 10, generator=gen).cuda(); hh0 = torch.randn(b2.num_nodes, 10, generator=gen).cuda()\n"
        "    out[name] = graph_of(b2, kind).layer_forward(kind, h, None if kind == N.KIND_DSS else hh0).cpu()\n"
        "torch.save(out, sys.argv[1])\n" % (ROOT, ROOT))
    outs = []
    for tag, env in (("fused", {}), ("prepass", {"PSI_NO_FUSED_PRE": "1"})):
        path = tmp_path / (tag + ".pt")
        e = dict(os.environ)
        e.pop("PSI_NO_FUSED_PRE", None)
        e.update(env)
        subprocess.run([sys.executable, str(script), str(path)], check=True, env=e, timeout=300)
        outs.append(torch.load(path))
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k


def test_node_permutation_equivariance():
    """relabelling the nodes permutes the output rows (the re-layout is independent of the input ordering)"""
    g = Golden("dirichlet_ckpt")
    m = g.model(DEV)
    b, h, h0, _ = _random_graph(5, 500, False, "cpu")
    perm = torch.randperm(500, generator=torch.Generator().manual_seed(1))
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(500)
    from psi_gnn_b200.synthetic import GraphData
    bp = GraphData(x=b.x[perm], edge_index=inv[b.edge_index], edge_attr=b.edge_attr, a_ij=b.a_ij, y=b.y[perm], sol=b.sol[perm],
                   prb_data=b.prb_data[perm], tags=b.tags[perm])
    bp.num_nodes = 500
    with torch.no_grad():
        a = m.deqdss.f(h.to(DEV), h0.to(DEV), b.to(DEV))
        c = m.deqdss.f(h[perm].to(DEV), h0[perm].to(DEV), bp.to(DEV))
    assert rel_err(c, a[perm.to(DEV)]) <= 1e-6


def test_solve_is_deterministic():
    """two runs of the whole fused Broyden solve give bit-identical iterates, traces and step counts (no atomics anywhere)"""
    g = Golden("mixed_ckpt")
    m = g.model(DEV)
    b = g.batch(DEV)
    h0 = g.t("h0", DEV)
    a = m.deqdss.inference(h0, b)
    c = m.deqdss.inference(h0, b)
    assert a["steps_run"] == c["steps_run"] and a["nstep"] == c["nstep"]
    assert a["rel_trace"] == c["rel_trace"]
    assert torch.equal(a["result"], c["result"])


def test_long_small_solves_are_bitwise_repeatable():
    """long Broyden solves (250 steps, no stop) of small meshes — 7 and 14 reduction chunks, where the dot pass packs its work items
    differently from the large configurations — repeated: identical residual traces to the last bit (fixed summation orders, no
    atomics, no timing dependence)"""
    from psi_gnn_b200 import solver as S, synthetic
    g = Golden("dirichlet_ckpt")
    m = g.model(DEV)
    for nodes in (3000, 6000):
        mesh = synthetic.make_large_mesh(nodes, seed=3).to(DEV)
        h0 = m._encode_native(mesh.x)
        op = S.LayerOperator(m.deqdss.f, h0, mesh)
        runs = [S.broyden(op, h0, threshold=250, eps=1e-30) for _ in range(5)]
        assert all(r["rel_trace"] == runs[0]["rel_trace"] for r in runs[1:]), nodes
        assert all(torch.equal(r["result"], runs[0]["result"]) for r in runs[1:]), nodes


def test_anderson_generic_callable():
    from psi_gnn_b200 import solver as S
    g = Golden("dirichlet_seed0")
    m = g.model(DEV)
    b = g.batch(DEV)
    h0 = g.t("h0", DEV)
    op = S.LayerOperator(m.deqdss.f, h0, b)
    out = S.anderson(lambda H: op(H), h0, m=2, threshold=60, eps=1e-4)
    ref = g["anderson_rel_trace"]
    assert len(out["rel_trace"]) == ref.shape[0]
    assert np.allclose(out["rel_trace"][:6], ref[:6], rtol=2e-2)
    assert rel_err(out["result"], g.t("anderson_result")) <= 1e-2


def test_unsupported_configurations_fail_loudly():
    from psi_gnn_b200.dirichlet.psignn import model as M
    from psi_gnn_b200.dirichlet.psignn.utilities import solver as S
    g = Golden("dirichlet_seed0")
    b = g.batch(DEV)
    cfg = g.cfg()
    cfg["solver"] = S.broyden
    cfg["n_layers"] = 2
    m2 = M.ModelDEQDSS(cfg).to(DEV)
    with pytest.raises(NotImplementedError):
        m2.inference(b)
    h0 = g.t("h0", DEV)
    with pytest.raises(NotImplementedError):
        S.broyden(lambda x: x, h0, threshold=3, eps=1e-3, ls=True)
    with pytest.raises(NotImplementedError):
        S.broyden(lambda x: x, h0, threshold=3, eps=1e-3, stop_mode="abs")
    with pytest.raises(RuntimeError):
        S.broyden(lambda x: x, h0.double(), threshold=3, eps=1e-3)


def test_solver_recovers_after_nonfinite_solve():
    """a solve that blows up (NaN input) stops early, returns its start point, and leaves nothing behind in the pooled workspace:
    the next solve on the same graph is bit-identical to one on a fresh workspace"""
    from psi_gnn_b200 import solver as S
    g = Golden("dirichlet_ckpt")
    m = g.model(DEV)
    b = g.batch(DEV)
    h0 = g.t("h0", DEV)
    op = S.LayerOperator(m.deqdss.f, h0, b)
    ref = S.broyden(op, h0, threshold=200, eps=1e-5)
    bad = h0.clone()
    bad[7, 3] = float("nan")
    out = S.broyden(op, bad, threshold=200, eps=1e-5)
    assert out["steps_run"] <= 2 and not (out["lowest"] < 1e-5)
    huge = h0 * 1e30
    S.broyden(op, huge, threshold=50, eps=1e-5)                 # overflows to inf within a step or two
    again = S.broyden(op, h0, threshold=200, eps=1e-5)
    assert again["steps_run"] == ref["steps_run"] and again["rel_trace"] == ref["rel_trace"]
    assert torch.equal(again["result"], ref["result"])


# ---- quasi-Newton kernels across sizes: a linear contraction is not chaotic, so the fp32 kernels must track the fp64 oracle ---------
@pytest.mark.parametrize("rows", [1, 7, 409, 410, 411, 1000, 4099, 20481])
def test_broyden_linear_operator_tracks_fp64_oracle(rows):
    """f(x) = d ⊙ x + 0.05·W(Wᵀx) + b with |d| < 0.8: Broyden converges in a few dozen steps; sizes straddle the 4096-float chunk,
    the 2048-float tile and the float4 granularities of the TMA kernels (rows × 10 floats), depths straddle the 32-vector work items."""
    from oracle import psignn_oracle as O
    from psi_gnn_b200 import solver as S
    gen = torch.Generator().manual_seed(rows)
    d = (torch.rand(rows, 10, generator=gen) * 1.6 - 0.8).double()
    Wm = (torch.randn(rows * 10, 3, generator=gen) / (rows * 10) ** 0.5).double()
    bvec = torch.randn(rows, 10, generator=gen).double()

    def f64(x):
        return d * x + 0.05 * (Wm @ (Wm.t() @ x.reshape(-1))).reshape(rows, 10) + bvec

    d32, W32, b32 = d.float().to(DEV), Wm.float().to(DEV), bvec.float().to(DEV)

    def f32(x):
        return d32 * x + 0.05 * (W32 @ (W32.t() @ x.reshape(-1))).reshape(rows, 10) + b32

    x0 = torch.zeros(rows, 10)
    thr = 80
    ref = O.broyden(f64, x0.double(), threshold=thr, eps=1e-6, keep_trace=False)
    out = S.broyden(f32, x0.to(DEV), threshold=thr, eps=1e-6)
    assert out["lowest"] < 1e-6 and ref["lowest"] < 1e-6
    assert abs(out["nstep"] - ref["nstep"]) <= 2
    k = min(out["steps_run"], len([r for r in ref["rel_trace"] if r > 1e-4]))
    got, want = np.asarray(out["rel_trace"][:k]), np.asarray(ref["rel_trace"][:k])
    assert np.all(np.abs(got - want) <= 2e-2 * want + 1e-7), (got, want)
    if rows <= 1000:
        xs = torch.linalg.solve(torch.eye(rows * 10, dtype=torch.float64) - torch.diag(d.reshape(-1)) - 0.05 * Wm @ Wm.t(), bvec.reshape(-1))
    else:
        xs = ref["result"].reshape(-1)
    assert float((out["result"].double().cpu().reshape(-1) - xs).norm() / xs.norm()) < 2e-5


# ---- BASELINE-configuration sizes against the reference + its fp64-tight truths --------------------------------
CFG_FIXTURES = ["cfg_c1", "cfg_c3shard", "cfg_c4mixed"]


@pytest.fixture(scope="module", params=CFG_FIXTURES)
def cfg_golden(request):
    import os
    from conftest import GOLDEN
    if not os.path.exists(os.path.join(GOLDEN, request.param + ".npz")):
        pytest.skip("fixture %s not generated" % request.param)
    return Golden(request.param)


def _stable_prefix(a, b, tol):
    """number of leading entries over which two rel traces agree to `tol` (relative)"""
    k = min(len(a), len(b))
    d = np.abs(np.asarray(a[:k]) - np.asarray(b[:k])) / np.asarray(b[:k])
    bad = np.nonzero(d > tol)[0]
    return int(bad[0]) if bad.size else k


def _edge_permuted(b, seed):
    from psi_gnn_b200.synthetic import GraphData
    perm = torch.randperm(b.edge_index.shape[1], generator=torch.Generator().manual_seed(seed)).to(b.edge_index.device)
    bp = GraphData()
    bp.__dict__.update({k: v for k, v in b.__dict__.items() if not k.startswith("_psi")})
    bp.edge_index, bp.edge_attr, bp.a_ij = b.edge_index[:, perm].contiguous(), b.edge_attr[perm].contiguous(), b.a_ij[perm].contiguous()
    return bp


def test_config_forward_solve(cfg_golden):
    """free-running forward solve at configuration size (32 ≈500-node meshes = C0/C1 and one 8-GPU shard of C3; 8 ≈2 k-node mixed
    meshes = C4 sample), protocol of SURVEY §8c (3).  A single free-running solve is one draw of a chaotic process — the fixtures
    hold ten runs of the REFERENCE ITSELF (edge list permuted: same graph, same weights), whose step counts span 74…102 and whose
    distances to the fp64-tight fixed point span 1.0e-3…2.2e-3 on cfg_c1 — so the CUDA path is sampled the same way (5 edge
    permutations) and held to the reference's own spread:
      * the first 20 rel-trace entries within 1e-3 of the reference, over the prefix on which the reference agrees with ITSELF under an
        edge permutation to 3e-4 (measured: 10 steps; beyond it the secant updates have amplified rounding noise in the reference too);
      * converged like the reference; every step count within ±5 % / ±3 of the interval spanned by the reference's runs;
      * ‖u − u_fp64,tight‖/‖u‖: median over our runs ≤ 1.25 × the largest deviation of the reference's runs, no run beyond 2 ×;
      * physics residual within the same band."""
    g = cfg_golden
    m = g.model(DEV)
    b = g.batch(DEV)
    h0 = g.t("h0", DEV)
    out = m.deqdss.inference(h0, b)
    ref_rel, perm_rel = g["fw_rel_trace"], g["perm_fw_rel_trace"]
    ref_run = int(g["fw_steps_run"])
    stable = _stable_prefix(perm_rel[:ref_run], ref_rel[:ref_run], 3e-4)
    k = min(20, stable, out["steps_run"])
    assert k >= 8, "the reference itself is not permutation-stable over a meaningful prefix (%d)" % stable
    got = np.asarray(out["rel_trace"][:k])
    assert np.all(np.abs(got - ref_rel[:k]) <= 1e-3 * ref_rel[:k]), (k, got, ref_rel[:k])
    eps = float(g["cfg.fw_tol"])
    ref_conv = float(g["fw_lowest"]) < eps
    assert bool(out["prot_break"]) == bool(int(g["fw_prot_break"]))
    # the reference's own runs
    ref_steps = [int(g["fw_nstep"]), int(g["perm_fw_nstep"])] + ([int(x) for x in g["spread_nstep"]] if g.has("spread_nstep") else [])
    ref_devs = [rel_err(g.t("u"), g.t("u64")), rel_err(g.t("perm_u"), g.t("u64"))] + ([float(x) for x in g["spread_u_dev64"]] if g.has("spread_u_dev64") else [])
    lo, hi = min(ref_steps), max(ref_steps)
    tol = max(3, int(round(0.05 * int(g["fw_nstep"]))))
    # ours: the stored edge order + 4 permutations
    runs = [(out, b)] + [(None, _edge_permuted(b, 3000 + i)) for i in range(4)]
    devs, steps, resid = [], [], []
    r64 = m.residual_loss(g.t("u64", DEV).float(), b).item()
    for o, bb in runs:
        o = o if o is not None else m.deqdss.inference(h0, bb)
        u = m._decode_native(o["result"])
        devs.append(rel_err(u, g.t("u64")))
        steps.append(o["nstep"])
        resid.append(abs(m.residual_loss(u, b).item() - r64))
        if ref_conv:
            assert o["lowest"] < eps
    print("config %s: nstep %s (reference runs %s, allowed [%d, %d]); u vs fp64-tight truth %s (reference runs %.2e … %.2e); rel trace "
          "checked over %d steps (reference self-stable over %d)" % (g.name, steps, sorted(ref_steps), lo - tol, hi + tol,
                                                                    " ".join("%.2e" % d for d in devs), min(ref_devs), max(ref_devs), k, stable))
    assert float(np.median(devs)) <= 1.25 * max(ref_devs), (devs, ref_devs)
    assert max(devs) <= 2.0 * max(ref_devs), (devs, ref_devs)
    if ref_conv:
        assert all(lo - tol <= s_ <= hi + tol for s_ in steps), (steps, lo, hi, tol)
    r_ref = max(abs(float(g["residual"]) - r64), abs(m.residual_loss(g.t("perm_u", DEV), b).item() - r64))
    assert float(np.median(resid)) <= 1.25 * r_ref + 1e-6 * abs(r64)


def _cfg_train(g, monkeypatch, pinned):
    from psi_gnn_b200 import solver as S
    m = g.model(DEV)
    b = g.batch(DEV)
    if pinned:
        hstar = g.t("train_hstar", DEV)
        calls = []

        def solver(f, x0, threshold, eps):
            if not calls:
                calls.append("fw")
                return {"result": hstar.clone(), "lowest": float(g["fw_lowest"]), "nstep": int(g["fw_nstep"]), "steps_run": 0, "f_evals": 0,
                        "launches": 0}
            return S.broyden(f, x0, threshold=threshold, eps=eps)

        m.deqdss.config_deq["solver"] = solver
    u, ld, gs = _train_step(m, b, g.t("train_v", DEV), monkeypatch)
    names = [k for k, _ in m.named_parameters()]
    cat = lambda pre: torch.cat([g.t(pre + k).reshape(-1).double() for k in names])
    return m, u, ld, gs, cat


def test_config_training_step_teacher_forced_vs_fp64_truth(cfg_golden, monkeypatch):
    """parameter gradients against the fp64 truth.  Forward fixed point pinned to the reference's fp32 H*; the truth is
    the reference's own code run in fp64 at that H* with the backward solve tightened to 1e-12.  The CUDA path (native VJP + Broyden
    backward solve + parameter gradients) must be as close to the truth as the reference's fp32 run is (×1.25, floor 1e-5)."""
    g = cfg_golden
    if not g.has("tf64_grad.deqdss.f.laynorm.weight"):
        pytest.skip("fixture holds no training step")
    m, u, ld, gs, cat = _cfg_train(g, monkeypatch, pinned=True)
    t64, r32 = cat("tf64_grad."), cat("train_grad.")
    e_ours = float((gs - t64).norm() / t64.norm())
    e_ref = float((r32 - t64).norm() / t64.norm())
    print("config %s teacher-forced gradient vs fp64 truth: ours %.3e, reference fp32 %.3e; backward solve lowest %.2e (reference %.2e)" % (
        g.name, e_ours, e_ref, m.deqdss.last_backward["lowest"], float(g["train_bw_lowest"])))
    assert e_ours <= max(1.25 * e_ref, 1e-5), (e_ours, e_ref)
    assert rel_err(u.detach(), g.t("train_u")) <= 1e-5
    for k in ("residual_loss", "jacobian_loss", "encoder_loss", "autoencoder_loss", "mse_loss", "mse_dirichlet"):
        ref = float(g["tf64_loss." + k])
        assert abs(ld[k].item() - ref) <= 1e-4 * abs(ref) + 1e-9, k


def test_config_training_step_free_running_vs_fp64_truth(cfg_golden, monkeypatch):
    """the same with the native forward solve, against the all-fp64 truth (both solves tight): the fixed-point error of any fp32
    forward solve (≈ 1e-3 in u) dominates here — for the reference as for us — so the band is the reference's own distance ×1.25"""
    g = cfg_golden
    if not g.has("train64_grad.deqdss.f.laynorm.weight"):
        pytest.skip("fixture holds no training step")
    m, u, ld, gs, cat = _cfg_train(g, monkeypatch, pinned=False)
    t64, r32 = cat("train64_grad."), cat("train_grad.")
    e_ours = float((gs - t64).norm() / t64.norm())
    e_ref = float((r32 - t64).norm() / t64.norm())
    print("config %s free-running gradient vs fp64 truth: ours %.3e, reference fp32 %.3e" % (g.name, e_ours, e_ref))
    assert torch.isfinite(gs).all()
    assert e_ours <= 1.25 * e_ref + 1e-5, (e_ours, e_ref)
    # stop behaviour of the backward solve (the reference runs into the bw_thres cap on trained weights: SURVEY §3.2)
    bw = m.deqdss.last_backward
    ref_steps, thr = int(g["train_bw_steps_run"]), int(g["cfg.bw_thres"])
    if ref_steps >= thr:
        assert bw["steps_run"] >= 0.8 * thr or bw["lowest"] < float(g["cfg.bw_tol"])
