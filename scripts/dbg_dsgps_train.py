"""diagnosis: parameter gradients of the unrolled DSGPS training step — native layer backward vs torch autograd (fp32, fp64) vs golden"""
import sys, os, copy
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import test_gpu_parity as T
from psi_gnn_b200 import baselines as B

name = sys.argv[1] if len(sys.argv) > 1 else "dsgps_ckpt"
g, m, b = T._baseline(name)


def grads(model, batch):
    model.zero_grad()
    U, ld = model(batch)
    grads.U = {k: v.detach().double().cpu() for k, v in U.items()}
    ld["train_loss"].backward()
    return {n: (p.grad.detach().double().cpu() if p.grad is not None else torch.zeros_like(p).double().cpu()) for n, p in model.named_parameters()}, float(ld["train_loss"].detach())


YB = {}


def hooked(inner, tag):
    class H:
        @staticmethod
        def apply(*a):
            out = inner.apply(*a)
            lst = YB.setdefault(tag, [])
            idx = len(lst)
            lst.append(None)
            out.register_hook(lambda gr, i=idx: lst.__setitem__(i, gr.detach().double().cpu()))
            return out
    return H


_native_layer = B._UnrolledLayer
B._UnrolledLayer = hooked(_native_layer, "native")
gn, ln = grads(m, b)
B._UnrolledLayer = _native_layer
Un = grads.U


class TorchLayer:
    """route _UnrolledLayer through the differentiable torch step"""
    @staticmethod
    def apply(owner, batch, step, block, dmask, h, h0, *params):
        return TorchLayer.model._step_torch(step, h, h0, TorchLayer.batch)


class HybridLayer(torch.autograd.Function):
    """native forward, torch-autograd backward of the same step recomputed at the saved input"""
    @staticmethod
    def forward(ctx, owner, batch, step, block, dmask, h, h0, *params):
        from psi_gnn_b200 import weights as W
        from psi_gnn_b200.graph import graph_of
        W.upload(*block)
        ctx.save_for_backward(h, h0)
        ctx.step, ctx.names = step, owner._layer_names(step)
        return graph_of(batch, owner._layer_kind).layer_forward(owner._layer_kind, h.detach(), h0.detach())

    @staticmethod
    def backward(ctx, ybar):
        h, h0 = ctx.saved_tensors
        mm, bb = TorchLayer.model, TorchLayer.batch
        P = dict(mm.named_parameters())
        with torch.enable_grad():
            h_ = h.detach().requires_grad_()
            h0_ = h0.detach().requires_grad_()
            out = mm._step_torch(ctx.step, h_, h0_, bb)
            gr = torch.autograd.grad(out, [h_, h0_] + [P[n] for n in ctx.names], ybar, allow_unused=True)
        return (None,) * 5 + tuple(gr)


orig = B._UnrolledLayer
B._UnrolledLayer = HybridLayer
TorchLayer.model, TorchLayer.batch = m, b
gh, lh = grads(m, b)
B._UnrolledLayer = hooked(TorchLayer, "t32")
_graph_of = B.graph_of
B.graph_of = lambda batch, kind: _graph_of(batch, kind) if batch.edge_attr.dtype == torch.float32 else None
TorchLayer.model, TorchLayer.batch = m, b
g32, l32 = grads(m, b)
U32 = grads.U
m64 = copy.deepcopy(m).double()
b64 = copy.copy(b)
for k, v in list(b.__dict__.items()):
    if torch.is_tensor(v) and v.dtype == torch.float32:
        setattr(b64, k, v.double())
TorchLayer.model, TorchLayer.batch = m64, b64
B._UnrolledLayer = hooked(TorchLayer, "t64")
# the residual loss is native fp32: use a torch one for the fp64 run
m64.residual_loss = lambda u, batch: torch.mean((torch.zeros_like(u).index_add(0, batch.edge_index[0], batch.a_ij.reshape(-1, 1) * u[batch.edge_index[1]]) - batch.y) ** 2)
g64, l64 = grads(m64, b64)
U64 = grads.U
for k in ("1", "2", "3", "5", "10", "20", "30"):
    if k in U64:
        r = float(U64[k].norm())
        print("U[%s] native %.2e torch32 %.2e" % (k, float((Un[k] - U64[k]).norm()) / r, float((U32[k] - U64[k]).norm()) / r))
B._UnrolledLayer = orig
gg = {n: g.t("train_grad." + n).double() for n in gn}
for k in range(len(YB["t64"])):
    r = YB["t64"][k]
    en, e32 = YB["native"][k] - r, YB["t32"][k] - r
    rows = en.norm(dim=1) ** 2
    top = torch.topk(rows, 3)
    print("ybar[%2d] native %.2e torch32 %.2e | top-3 rows hold %.0f%% of the native error: %s" % (
        k + 1, float(en.norm() / r.norm()), float(e32.norm() / r.norm()), 100 * float(top.values.sum() / rows.sum()), top.indices.tolist()))
print("loss native %.8e torch32 %.8e torch64 %.8e golden %.8e" % (ln, l32, l64, float(g["train_loss"])))


def tot(a, ref):
    num = sum(float((a[n] - ref[n]).norm() ** 2) for n in a) ** 0.5
    den = sum(float(ref[n].norm() ** 2) for n in a) ** 0.5
    return num / den


print("hybrid (native fwd, torch bwd) vs fp64 %.3e, vs native %.3e" % (tot(gh, g64), tot(gh, gn)))
print("total vs fp64: native %.3e torch32 %.3e golden %.3e | native vs golden %.3e, torch32 vs golden %.3e" % (tot(gn, g64), tot(g32, g64), tot(gg, g64), tot(gn, gg), tot(g32, gg)))
for n in gn:
    r = float(g64[n].norm())
    if r == 0:
        continue
    print("%-40s |g| %.3e  native %.2e torch32 %.2e golden %.2e" % (n, r, float((gn[n] - g64[n]).norm()) / r, float((g32[n] - g64[n]).norm()) / r, float((gg[n] - g64[n]).norm()) / r))
