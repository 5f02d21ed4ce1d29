"""How good a residual does the teacher-forced backward solve reach at the 500-step cap?  (sampled over tiny rescalings of grad)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import Golden
from psi_gnn_b200 import solver as S
for name in ("mixed_ckpt",):
    g = Golden(name)
    m = g.model("cuda:0")
    b = g.batch("cuda:0")
    res = []
    for sc in [1.0 + 1e-4 * i for i in range(-16, 17)]:
        grad = g.t("train_bw_grad", "cuda:0") * sc
        op = S.VjpOperator(m.deqdss.f, g.t("train_hstar", "cuda:0"), b, grad)
        out = S.broyden(op, torch.zeros_like(grad), threshold=500, eps=1e-8)
        res.append("%.1e" % out["lowest"])
    print(os.environ.get("PSI_GNN_B200_LIB", "new")[-12:], name, "ref lowest %.1e" % float(g["train_bw_lowest"]), " ".join(res))
