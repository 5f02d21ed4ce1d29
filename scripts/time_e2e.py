"""Where does the end-to-end step go?  Times H2D, graph re-layout, forward solve, f(H*)+losses, backward on C3."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from psi_gnn_b200 import model as PM
from psi_gnn_b200.dirichlet.psignn import model as M
from psi_gnn_b200.dirichlet.psignn.utilities import solver as S

dev = torch.device("cuda:0")
P, cfg = bench.load_params("dirichlet")
cfg["solver"] = S.broyden
m = M.ModelDEQDSS(cfg); m.load_state_dict(P); m = m.to(dev).train()
hb = bench.make_batch("dirichlet", 256, 0.075, 0).pin_memory()


def T():
    torch.cuda.synchronize(); return time.perf_counter()


for it in range(4):
    t0 = T(); b = hb.to(dev, non_blocking=True)
    t1 = T(); g = PM.graph_of(b, 0)
    t2 = T(); h0 = m.autoencoder.encoder(b.x)
    with torch.no_grad():
        out = m.deqdss._solve(S.LayerOperator(m.deqdss.f, h0, b), h0, "fw_thres", "fw_tol")
    t3 = T()
    m.zero_grad()
    u, ld = m(b)
    t4 = T()
    loss = ld["residual_loss"] + ld["jacobian_loss"] + ld["encoder_loss"] + ld["autoencoder_loss"]
    loss.backward()
    t5 = T()
    del b, g, u, ld, loss, out
    t6 = T()
    print("it%d h2d %.1f ms | graph %.1f | fw solve %.1f (steps %d) | model fwd (incl. 2nd fw solve) %.1f | backward %.1f (bw steps %d) | free %.1f" % (
        it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), m.deqdss.last_forward["steps_run"], 1e3 * (t4 - t3), 1e3 * (t5 - t4),
        m.deqdss.last_backward["steps_run"], 1e3 * (t6 - t5)))
