"""CPU: host-side logic — the C-ABI library loads and exports every symbol the header declares, the weight packing
matches the struct layout, the drop-in modules keep the reference's state_dict keys, and nothing silently runs on CPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, Golden


def test_library_exports_every_declared_symbol():
    from psi_gnn_b200 import _native as N
    from psi_gnn_b200 import build as B
    B.build()
    lib = ctypes.CDLL(N.lib_path())
    header = open(os.path.join(ROOT, "include", "psignn_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(psi_[a-z0-9_]+)\s*\(", body))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(N.SIGNATURES), declared ^ set(N.SIGNATURES)
    assert lib.psi_version() >= 100


def test_weight_block_layout():
    from psi_gnn_b200 import _native as N
    from psi_gnn_b200 import weights as W
    assert N.load().psi_weights_floats() == W.TOTAL_FLOATS == 3200 + 2760      # struct LayerWeights + struct LayerWeightsT
    g = Golden("mixed_seed0")
    P = g.params()
    blob = W.pack_psignn(P, True, "cpu")
    o = W.OFFSETS
    W1 = P["deqdss.f.phi_from_list.0.mlp.mlp.0.weight"]
    assert torch.equal(blob[o["from.W1j"]:o["from.W1j"] + 100].view(10, 10), W1[:, 10:20])
    assert torch.equal(blob[o["from.W1a"]:o["from.W1a"] + 30].view(10, 3), W1[:, 20:23])
    up = P["deqdss.f.update_list.0.mlp.0.weight"]
    assert torch.equal(blob[o["up_W1"]:o["up_W1"] + 330].view(10, 33), up)
    un = P["deqdss.f.update_neumann.mlp.0.weight"]
    assert torch.equal(blob[o["un_W1"]:o["un_W1"] + 250].view(10, 25), un)
    assert torch.equal(blob[o["dec_W2"]:o["dec_W2"] + 10], P["autoencoder.decoder.mlp.mlp.2.weight"].reshape(-1))
    # the transposed tail (packed-FMA kernels): [input][output] copies of the same matrices
    t0 = W.MAIN_FLOATS
    assert torch.equal(blob[t0:t0 + 100].view(10, 10), P["deqdss.f.phi_to_list.0.mlp.mlp.0.weight"][:, 0:10].t())
    upT = t0 + 3 * 330
    assert torch.equal(blob[upT:upT + 330].view(33, 10), up.t())
    # dirichlet: prb width 2 → the 33rd column of the gate/update rows stays zero
    gd = Golden("dirichlet_seed0")
    bd = W.pack_psignn(gd.params(), False, "cpu")
    assert float(bd[o["up_W1"]:o["up_W1"] + 330].view(10, 33)[:, 32].abs().max()) == 0.0
    assert float(bd[o["neu.W1i"]:o["gate_w"]].abs().max()) == 0.0


def test_state_dict_keys_match_reference_checkpoints(golden):
    m = golden.model("cpu")
    assert set(m.state_dict().keys()) == set(golden.params().keys())
    n_params = sum(p.numel() for p in m.parameters())
    assert n_params == (2175 if golden.mixed else 1444)       # reference logs/model_config.csv


def test_constructor_signatures():
    import inspect
    from psi_gnn_b200.dirichlet.psignn import model as M
    from psi_gnn_b200.dirichlet.psignn.utilities import solver as S
    assert list(inspect.signature(M.Function.__init__).parameters)[1:] == ["n_layers", "latent_dim", "edge_features_dim",
                                                                            "second_member_dim", "activation"]
    assert list(inspect.signature(M.DeepEquilibrium.__init__).parameters)[1:] == ["function", "config_deq"]
    assert list(inspect.signature(S.broyden).parameters)[:7] == ["f", "x0", "threshold", "eps", "stop_mode", "ls", "name"]
    assert list(inspect.signature(S.anderson).parameters)[:8] == ["f", "x0", "m", "lam", "threshold", "eps", "stop_mode", "beta"]
    assert list(inspect.signature(S.forward_iteration).parameters)[:4] == ["f", "z0", "eps", "threshold"]      # + keep_trace, **kwargs (accepted, as the drop-in DeepEquilibrium.inference forwards keep_trace to every solver)


def test_no_cpu_fallback(golden):
    m = golden.model("cpu")
    b = golden.batch("cpu")
    h0 = golden.t("h0")
    with pytest.raises(RuntimeError):
        m.inference(b)
    with pytest.raises(RuntimeError):
        m.deqdss.f(h0, h0, b)
    with pytest.raises(RuntimeError):
        m.residual_loss(golden.t("u"), b)
    from psi_gnn_b200 import solver as S
    with pytest.raises(RuntimeError):
        S.broyden(lambda x: x, h0, threshold=3, eps=1e-3)
    with pytest.raises(RuntimeError):
        S.forward_iteration(lambda x: x, h0)


def test_product_never_imports_oracle():
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import psi_gnn_b200.model, psi_gnn_b200.solver, psi_gnn_b200.graph, "
            "psi_gnn_b200.dirichlet.psignn.model, psi_gnn_b200.mixed.psignn.model; "
            "assert not any(k == 'oracle' or k.startswith('oracle.') for k in sys.modules), 'oracle imported'") % ROOT
    subprocess.run([sys.executable, "-c", code], check=True)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "psi_gnn_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_synthetic_generator_contract():
    from psi_gnn_b200 import synthetic
    b = synthetic.make_batch(2, seed0=3, h=0.15)
    n, nnz = b.num_nodes, b.edge_index.shape[1]
    assert b.x.shape == (n, 1) and b.y.shape == (n, 1) and b.prb_data.shape == (n, 2) and b.tags.shape == (n, 1)
    assert b.edge_attr.shape == (nnz, 3) and b.a_ij.shape == (nnz, 1) and b.edge_index.dtype == torch.int64
    # Dirichlet rows of A hold only the diagonal (bc.apply), so A u = b there means u = g
    row, col = b.edge_index
    d = torch.where(b.tags.reshape(-1) == 1)[0]
    for i in d[:5].tolist():
        sel = row == i
        assert int(sel.sum()) == 1 and int(col[sel][0]) == i and float(b.a_ij[sel][0]) == 1.0
    parts = synthetic.split_graphs(b, 2)
    assert sum(p.num_nodes for p in parts) == n
    m = synthetic.make_batch(1, seed0=3, h=0.15, mixed=True)
    assert m.tags.shape[1] == 3 and m.prb_data.shape[1] == 3 and m.unit_normal_vector.shape == (m.num_nodes, 2)
    assert float(m.tags.sum(1).min()) == 1.0 and float(m.tags.sum(1).max()) == 1.0


def test_weight_block_layout_baselines():
    from psi_gnn_b200 import weights as W
    o = W.OFFSETS
    gd = Golden("dss_ckpt")
    P = gd.params()
    k = 7
    blob = W.pack_dss(P, k, 1e-3, "cpu")
    W1 = P["phi_to_list.%d.mlp.mlp.0.weight" % k]                      # [10, 21]: h_i, h_j, a_ij
    assert torch.equal(blob[o["to.W1a"]:o["to.W1a"] + 30].view(10, 3)[:, 0], W1[:, 20])
    assert float(blob[o["to.W1a"]:o["to.W1a"] + 30].view(10, 3)[:, 1:].abs().max()) == 0.0
    assert torch.equal(blob[o["up_W1"]:o["up_W1"] + 330].view(10, 33), P["psi_list.%d.mlp.mlp.0.weight" % k])
    assert torch.equal(blob[o["dec_W1"]:o["dec_W1"] + 100].view(10, 10), P["decoder_list.%d.mlp.mlp.0.weight" % k])
    assert abs(float(blob[o["dss_alpha"]]) - 1e-3) < 1e-9               # fp32(1e-3), what `alpha * tensor` uses in the reference too
    gg = Golden("dsgps_ckpt")
    Pg = gg.params()
    bg = W.pack_dsgps(Pg, "cpu")
    assert torch.equal(bg[o["gz_W"]:o["gz_W"] + 330].view(10, 33)[:, :32], Pg["z_k.mlp.0.weight"])        # row pitch 33: the mixed family has 33 inputs
    assert torch.equal(bg[o["gc_b"]:o["gc_b"] + 10], Pg["correction.mlp.0.bias"])
    assert torch.equal(bg[o["enc_W1"]:o["enc_W1"] + 10], Pg["autoencoder.encoder.mlp.mlp.0.weight"].reshape(-1))


def test_baseline_state_dict_keys():
    from psi_gnn_b200.dirichlet.dss import model as DSS
    from psi_gnn_b200.dirichlet.dsgps import model as DSGPS
    gd, gg = Golden("dss_ckpt"), Golden("dsgps_ckpt")
    m = DSS.DeepStatisticalSolver(dict(latent_dim=10, k=int(gd["cfg.k"]), alpha=float(gd["cfg.alpha"]), gamma=0.9))
    assert set(m.state_dict().keys()) == set(gd.params().keys())          # 480 tensors of the shipped DSS checkpoint
    m2 = DSGPS.ModelDSGPS(dict(latent_dim=10, k=30, alpha=1e-3, gamma=0.9))
    assert set(m2.state_dict().keys()) == set(gg.params().keys())
    with pytest.raises(RuntimeError):                                     # the training forward exists, but there is no CPU path
        m.forward(gd.batch("cpu"))


def test_shard_batch_balanced_and_complete():
    from psi_gnn_b200 import parallel, synthetic
    b = synthetic.make_batch(7, seed0=50, h=0.2)
    for world in (1, 2, 3, 7):
        shards = [parallel.shard_batch(b, r, world) for r in range(world)]
        assert sum(s.num_graphs for s in shards) == 7 and sum(s.num_nodes for s in shards) == b.num_nodes
        x = torch.cat([s.x for s in shards])
        assert torch.equal(x, b.x)
        for s in shards:
            assert s.num_graphs >= 1 and int(s.edge_index.max()) < s.num_nodes
    with pytest.raises(ValueError):
        parallel.shard_batch(b, 0, 8)


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm the driver runs beside the native arm): exactly one stdout line, the contract's keys"""
    import json
    import subprocess
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_loader_and_batch_standins():
    """DataListLoader / Batch.from_data_list (PyG stand-ins of psi_gnn_b200/training.py): list batches, rank shares, offsets"""
    from psi_gnn_b200 import synthetic, training as T
    data = [synthetic.make_mesh(300 + i, h=0.2) for i in range(7)]
    loader = T.DataListLoader(data, batch_size=3)
    batches = list(loader)
    assert [len(b) for b in batches] == [3, 3, 1] and len(loader) == 3
    b = T.Batch.from_data_list(batches[0])
    assert b.num_graphs == 3 and b.num_nodes == sum(d.num_nodes for d in batches[0]) and int(b.ptr[-1]) == b.num_nodes
    assert int(b.edge_index.max()) < b.num_nodes and int(b.edge_index[:, int(b.edge_ptr[1]):].min()) >= int(b.ptr[1])
    r0 = list(T.DataListLoader(data, batch_size=4, rank=0, world=2))
    r1 = list(T.DataListLoader(data, batch_size=4, rank=1, world=2))
    assert [len(x) for x in r0] == [2, 2] and [len(x) for x in r1] == [2, 1]
    assert [d._psi_item_id for d in r0[0]] == [0, 1] and [d._psi_item_id for d in r1[0]] == [2, 3]
    sh = T.DataListLoader(data, batch_size=7, shuffle=True, seed=1)
    e1 = [d._psi_item_id for d in next(iter(sh))]
    e2 = [d._psi_item_id for d in next(iter(sh))]
    assert sorted(e1) == list(range(7)) and e1 != e2


def test_mixed_dsgps_state_dict_keys():
    from conftest import Golden
    from psi_gnn_b200.mixed.dsgps import model as M
    g = Golden("dsgps_mixed_ckpt")
    m = M.ModelDSGPS(dict(latent_dim=10, k=30, alpha=1e-3, gamma=0.9))
    assert set(m.state_dict().keys()) == set(g.params().keys())
    m.load_state_dict(g.params())
