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


def test_mesh_partition_reproduces_global_layer():
    """host logic of the mesh partition: on every rank's local graph (owned + ghost rows) the layer evaluated by the oracle
    equals the global evaluation on the owned rows, and the exchange lists are mutually consistent."""
    import numpy as np
    from conftest import Golden
    from oracle import psignn_oracle as O
    from psi_gnn_b200 import partition, synthetic
    g = Golden("dirichlet_ckpt")
    P = g.params()
    mesh = synthetic.make_mesh(5, h=0.06)
    gen = torch.Generator().manual_seed(0)
    h = torch.randn(mesh.num_nodes, 10, generator=gen)
    h0 = torch.randn(mesh.num_nodes, 10, generator=gen)
    with torch.no_grad():
        ref = O.f_dirichlet(P, h, h0, mesh)
        ref_res = O.residual_vector(mesh.sol, mesh)
    world = 3
    parts = partition.partition_mesh(mesh, world)
    assert sorted(np.concatenate([p.owned_global for p in parts]).tolist()) == list(range(mesh.num_nodes))
    for p in parts:
        loc = p.local
        nodes = torch.from_numpy(np.concatenate([p.owned_global, p.ghost_global]))
        assert loc.num_nodes == p.n_owned + p.n_ghost == nodes.numel()
        with torch.no_grad():
            out = O.f_dirichlet(P, h[nodes], h0[nodes], loc)
            res = O.residual_vector(mesh.sol[nodes], loc)
        own = torch.from_numpy(p.owned_global)
        assert float((out[:p.n_owned] - ref[own]).abs().max()) <= 1e-5
        assert float((res[:p.n_owned] - ref_res[own]).abs().max()) <= 1e-4 * float(ref_res.abs().max())
        # exchange lists: what I send to q is exactly q's ghost segment owned by me, in q's order
        off = 0
        for q, cnt in zip(p.peers, p.send_counts):
            mine = p.owned_global[p.send_index[off:off + cnt]]
            off += cnt
            other = parts[q]
            roff = 0
            for qq, rc in zip(other.peers, other.recv_counts):
                if qq == p.rank:
                    assert rc == cnt and np.array_equal(other.ghost_global[roff:roff + rc], mine)
                roff += rc
    single = partition.partition_mesh(mesh, world, rank=1)[0]
    assert np.array_equal(single.owned_global, parts[1].owned_global) and np.array_equal(single.send_index, parts[1].send_index)


def test_mesh_partition_aligned_cuts():
    """large enough meshes are cut at multiples of 2048 nodes (what makes the partitioned solve retrace the single-GPU trajectory),
    and the 1-rank 'partition' is a pure reordering of the mesh"""
    import numpy as np
    from conftest import Golden
    from oracle import psignn_oracle as O
    from psi_gnn_b200 import partition, synthetic
    mesh = synthetic.make_large_mesh(9500, seed=1)
    assert mesh.num_nodes >= 2 * partition.ALIGN_NODES * 2
    parts = partition.partition_mesh(mesh, 2)
    assert parts[0].n_owned % partition.ALIGN_NODES == 0 and parts[0].n_owned + parts[1].n_owned == mesh.num_nodes
    assert abs(parts[0].n_owned - mesh.num_nodes / 2) <= partition.ALIGN_NODES
    # geometric order: the ranks own horizontal strips — rank 0 lies below rank 1 up to the one lattice row the cut runs through
    y0 = mesh.pos[torch.from_numpy(parts[0].owned_global), 1]
    y1 = mesh.pos[torch.from_numpy(parts[1].owned_global), 1]
    spacing = float(((mesh.pos[:, 0].max() - mesh.pos[:, 0].min()) * (mesh.pos[:, 1].max() - mesh.pos[:, 1].min()) / mesh.num_nodes) ** 0.5)
    assert float(y0.max()) <= float(y1.min()) + spacing
    assert parts[0].n_ghost < 8 * mesh.num_nodes ** 0.5 and parts[0].peers == [1]
    one = partition.reorder_mesh(mesh)
    P = Golden("dirichlet_ckpt").params()
    gen = torch.Generator().manual_seed(2)
    h = torch.randn(mesh.num_nodes, 10, generator=gen)
    h0 = torch.randn(mesh.num_nodes, 10, generator=gen)
    order = torch.from_numpy(one.partition.owned_global)
    with torch.no_grad():
        a = O.f_dirichlet(P, h, h0, mesh)
        b = O.f_dirichlet(P, h[order], h0[order], one)
    assert one.partition.n_ghost == 0 and float((b - a[order]).abs().max()) <= 1e-5
    # ranks of the aligned partition own contiguous slices of that order
    assert np.array_equal(np.concatenate([p.owned_global for p in parts]), one.partition.owned_global)
