"""Where does the end-to-end overhead of a new batch go?  H2D copy and psi_graph_create for C3 (256 meshes) and C5 (one 1M-node mesh).
    python scripts/time_graph_create.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from psi_gnn_b200 import model as PM, partition, synthetic

dev = torch.device("cuda:0")


def T():
    torch.cuda.synchronize(); return time.perf_counter()


for name, host in (("c3", bench.make_batch("dirichlet", 256, 0.075, 0).pin_memory()),
                   ("c5", partition.reorder_mesh(synthetic.make_large_mesh(1_000_000, seed=0)).pin_memory())):
    for it in range(4):
        t0 = T(); b = host.to(dev, non_blocking=True)
        t1 = T(); g = PM.graph_of(b, 0)
        t2 = T(); del g, b
        t3 = T()
        print("%s it%d: N=%d nnz=%d | h2d %.2f ms | graph_of %.2f ms | free %.2f ms" % (name, it, host.num_nodes, host.edge_index.shape[1],
                                                                                         1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
