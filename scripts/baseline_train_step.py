"""one training step (forward + backward) of each baseline on its golden fixture — used for the ncu launch list of the native
unrolled-layer path (profiles/r02_baseline_launches.md) and as a timing probe"""
import os
import sys
import time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import test_gpu_parity as T

names = sys.argv[1:] or ["dss_ckpt", "dsgps_ckpt", "dsgps_mixed_ckpt"]
for name in names:
    g, m, b = T._baseline(name)
    m.train()
    for it in range(3):
        m.zero_grad()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        U, ld = m(b)
        ld["train_loss"].backward()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print("%s: training step %.2f ms (k = %d unrolled steps, %d nodes), loss %.6e" % (name, 1e3 * dt, m.config["k"], b.num_nodes, float(ld["train_loss"].detach())))
