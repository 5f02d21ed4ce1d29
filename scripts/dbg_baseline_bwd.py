"""debug helper: one native backward of a mixed DSGPS step (run under compute-sanitizer)"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import test_gpu_parity as T
from psi_gnn_b200 import _native as N, weights as W
from psi_gnn_b200.graph import graph_of
name = sys.argv[1] if len(sys.argv) > 1 else "dsgps_mixed_ckpt"
g, m, b = T._baseline(name)
kind = N.KIND_DSGPS_MIXED if "mixed" in name else (N.KIND_DSS if name.startswith("dss") else N.KIND_DSGPS)
n = b.num_nodes
h = torch.randn(n, 10, device="cuda") * 0.5
y = torch.randn(n, 10, device="cuda")
if kind == N.KIND_DSS:
    W.upload(*m._layer_block(3, "cuda:0"))
else:
    W.upload(*m._layer_block(0, torch.device("cuda:0")))
gr = graph_of(b, kind)
hb, flat = gr.layer_backward(kind, h, y)
torch.cuda.synchronize()
print("ok", float(hb.norm()), float(flat.norm()))
