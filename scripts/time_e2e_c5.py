"""Where does the end-to-end time of a C5 forward solve on a NEW device batch go?  (bench.py's e2e closure, piece by piece)
    python scripts/time_e2e_c5.py [nodes]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from psi_gnn_b200 import model as PM, partition, synthetic
from psi_gnn_b200.dirichlet.psignn import model as M
from psi_gnn_b200.dirichlet.psignn.utilities import solver as S

dev = torch.device("cuda:0")
nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
P, cfg = bench.load_params("dirichlet")
cfg["solver"] = S.broyden
m = M.ModelDEQDSS(cfg); m.load_state_dict(P); m = m.to(dev).eval()
host = partition.reorder_mesh(synthetic.make_large_mesh(nodes, seed=0)).pin_memory()
dev_batch = host.to(dev)


def T():
    torch.cuda.synchronize(); return time.perf_counter()


with torch.no_grad():
    for it in range(2):
        t0 = T(); m.inference(dev_batch); t1 = T()
        print("resident batch: solve %.1f ms (steps %d)" % (1e3 * (t1 - t0), m.deqdss.last_forward["steps_run"]))
    for it in range(3):
        t0 = T(); b = host.to(dev, non_blocking=True)
        t1 = T(); g = PM.graph_of(b, 0)
        t2 = T(); h0 = m._encode_native(b.x)
        t3 = T(); out = m.deqdss.inference(h0, b)
        t4 = T(); u = m._decode_native(out["result"]).cpu()
        t5 = T(); del b, g, h0, out, u
        t6 = T()
        print("new batch it%d: h2d %.1f | graph %.1f | encode %.1f | solve %.1f (steps %d) | decode + d2h %.1f | free %.1f ms" % (
            it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), m.deqdss.last_forward["steps_run"], 1e3 * (t5 - t4), 1e3 * (t6 - t5)))
