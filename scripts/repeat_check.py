"""Bitwise repeatability of long single-GPU Broyden solves at small sizes (diagnosis of the partitioned-solve glitch):
    python scripts/repeat_check.py [nodes ...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Golden
from psi_gnn_b200 import synthetic, solver as S
g = Golden("dirichlet_ckpt")
m = g.model("cuda:0")
for nodes in [int(a) for a in sys.argv[1:]] or [6200, 11400]:
    mesh = synthetic.make_large_mesh(nodes, seed=3).to("cuda:0")
    h0 = m._encode_native(mesh.x)
    op = S.LayerOperator(m.deqdss.f, h0, mesh)
    tr = [S.broyden(op, h0, threshold=300, eps=1e-30)["rel_trace"][:300] for _ in range(int(os.environ.get('PSI_REPEATS', '8')))]
    firsts = [next((i for i, (a, b) in enumerate(zip(t, tr[0])) if a != b), -1) for t in tr]
    print("N=%d numel=%d: first differing step vs run 0: %s" % (mesh.num_nodes, mesh.num_nodes * 10, firsts))
