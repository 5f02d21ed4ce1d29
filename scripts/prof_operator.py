"""Operator kernels in isolation (for ncu and quick CUDA-event timing): the fused layer and the VJP pair at C3 / C5 size.

    python scripts/prof_operator.py [c3|c5|c0] [reps]

Prints per-launch CUDA-event times (L2 flushed between launches) and the algorithmic GB/s (SURVEY §8d byte formulas as counted by
``operator_bytes`` in csrc/psignn_b200.cu, without the solver epilogue)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from conftest import Golden
    from psi_gnn_b200 import partition, synthetic
    from psi_gnn_b200.solver import VjpOperator
    which = sys.argv[1] if len(sys.argv) > 1 else "c5"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    mixed = which == "c4"
    dev = torch.device("cuda:0")
    g = Golden("mixed_ckpt" if mixed else "dirichlet_ckpt")
    m = g.model(dev)
    if which == "c5":
        mesh = partition.reorder_mesh(synthetic.make_large_mesh(int(os.environ.get("PSI_NODES", "1000000")), seed=0))
    elif which == "c4":
        mesh = synthetic.make_batch(256, seed0=0, h=0.037, mixed=True, solve=False)
    else:
        mesh = synthetic.make_batch(256 if which == "c3" else 32, seed0=0, solve=False)
    b = mesh.to(dev)
    f = m.deqdss.f
    N = b.num_nodes
    E = int((b.edge_index[0] != b.edge_index[1]).sum())
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    with torch.no_grad():
        h0 = m._encode_native(b.x)
        h = f(h0, h0, b)
        h = f(h, h0, b)
    y = torch.randn_like(h)
    op = VjpOperator(f, h, b, torch.zeros_like(h))
    if os.environ.get("PSI_CHECK_DET"):
        # bitwise run-to-run determinism at this size: layer, VJP, whole forward solve
        with torch.no_grad():
            a1, a2, a3 = f(h, h0, b), f(h, h0, b), f(h, h0, b)
        v1, v2 = op(y), op(y)
        s1 = m.deqdss.inference(h0, b)
        s2 = m.deqdss.inference(h0, b)
        s3 = m.deqdss.inference(h0, b)
        print("determinism %s: layer %s %s | vjp %s | solve steps %d %d %d, results equal %s %s, rel traces equal %s" % (
            which, torch.equal(a1, a2), torch.equal(a1, a3), torch.equal(v1, v2), s1["steps_run"], s2["steps_run"], s3["steps_run"],
            torch.equal(s1["result"], s2["result"]), torch.equal(s1["result"], s3["result"]), s1["rel_trace"] == s2["rel_trace"]))
        if s1["rel_trace"] != s2["rel_trace"]:
            d = [i for i, (p_, q_) in enumerate(zip(s1["rel_trace"], s2["rel_trace"])) if p_ != q_]
            print("   first differing step", d[0], s1["rel_trace"][d[0]], s2["rel_trace"][d[0]])

    def timed(fn):
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            c.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(c) * 1e3)
        return float(np.median(ts)), float(np.min(ts))

    lists = 3 if mixed else 2
    ndir = int(f.native_graph(b).num_dirichlet)
    bf = N * (40 + 40 + 1 + 4 * (3 if mixed else 2) + 0.5) + lists * E * 16 + ndir * 40
    bv = N * (253 + 120) + 2 * E * 8 + N * (80 + 40 + 40 + 40)
    with torch.no_grad():
        t_f = timed(lambda: f(h, h0, b))
    t_v = timed(lambda: op(y))
    print("%s: N=%d E=%d | layer %.1f us median (%.1f min) = %.0f GB/s algorithmic | vjp pair %.1f us (%.1f min) = %.0f GB/s" % (
        which, N, E, t_f[0], t_f[1], bf / t_f[0] / 1e3, t_v[0], t_v[1], bv / t_v[0] / 1e3))


if __name__ == "__main__":
    main()
