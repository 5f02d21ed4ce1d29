"""CPU, world_size 2 over gloo: the host logic of the graph-sharded multi-GPU path (psi_gnn_b200/parallel.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from psi_gnn_b200 import parallel, synthetic
    from psi_gnn_b200.dirichlet.psignn import model as M
    from psi_gnn_b200.dirichlet.psignn.utilities import solver as S
    try:
        # every rank builds the same global batch and takes its own shard
        batch = synthetic.make_batch(5, seed0=0, h=0.2)
        shard = parallel.shard_batch(batch, rank, world)
        n_all = torch.tensor([shard.num_nodes, shard.num_graphs], dtype=torch.int64)
        dist.all_reduce(n_all)
        assert int(n_all[0]) == batch.num_nodes and int(n_all[1]) == batch.num_graphs
        assert int(shard.edge_index.max()) < shard.num_nodes and int(shard.edge_index.min()) >= 0
        # identical replicas, rank-dependent gradients → averaged gradients everywhere
        torch.manual_seed(0)
        cfg = dict(latent_dim=10, hidden_dim=10, n_layers=1, fw_tol=1e-5, fw_thres=10, bw_tol=1e-8, bw_thres=10, solver=S.broyden, path_logs=None)
        m = M.ModelDEQDSS(cfg)
        params = list(m.parameters())
        for i, p in enumerate(params):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        params[3].grad = None if rank == 0 else params[3].grad        # a missing gradient counts as zero
        parallel.allreduce_gradients(params, world)
        for i, p in enumerate(params):
            want = (i + 1) * (1 + 2) / 2.0 if i != 3 else (i + 1) * 2 / 2.0
            assert torch.allclose(p.grad, torch.full_like(p, want)), (i, float(p.grad.flatten()[0]), want)
        vals = parallel.allreduce_scalars([float(rank), 1.0], "cpu")
        assert vals == [1.0, 2.0]
        q.put((rank, "ok"))
    except Exception as e:          # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_graph_sharded_gradient_allreduce_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
