"""One short Broyden solve of a small mesh (for compute-sanitizer --tool racecheck / memcheck):
    compute-sanitizer --tool racecheck python scripts/racecheck_small.py [nodes] [steps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Golden
from psi_gnn_b200 import synthetic, solver as S
nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 6200
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 230
g = Golden("dirichlet_ckpt")
m = g.model("cuda:0")
mesh = synthetic.make_large_mesh(nodes, seed=3).to("cuda:0")
h0 = m._encode_native(mesh.x)
op = S.LayerOperator(m.deqdss.f, h0, mesh)
out = S.broyden(op, h0, threshold=steps, eps=1e-30)
print("N=%d steps %d lowest %.3e" % (mesh.num_nodes, out["steps_run"], out["lowest"]))
