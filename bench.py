#!/usr/bin/env python
"""bench.py — PSI-GNN training-step hot path (forward Broyden solve + implicit-adjoint backward solve) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload c3|c4|c1]

One step = one training step of ``ModelDEQDSS`` on one batch through the drop-in module API: encoder → no-grad forward
fixed-point solve → differentiable f(H*) + Jacobian regulariser → decoder → losses → ``loss.backward()`` (implicit
backward solve inside the hook) → gradient all-reduce (N>1) → clip → two Adam steps.  Default workload = BASELINE.json
config 3 (256 synthetic ~500-node Poisson meshes per GPU, Dirichlet), weights = the reference's shipped checkpoint
(tests/golden/dirichlet_ckpt.npz).  Prints ONE JSON line (see README / DESIGN.md §6 for the keys).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

warnings.filterwarnings("ignore")
import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (family, graphs per GPU, mesh size h, description)
    "c0": ("dirichlet", 32, 0.075, "C0 (BASELINE configs[0]): PSI-GNN dirichlet forward fixed-point solve (inference: encoder, Broyden, decoder), batch of 32 "
                                    "synthetic ~500-node 2D triangle Poisson meshes per GPU"),
    "c2": ("dss", 32, 0.075, "C2 (BASELINE configs[1]): DSS baseline inference (30 unrolled layers with per-layer weights on the shared fused layer kernel), "
                              "batch of 32 synthetic ~500-node meshes per GPU, shipped DSS checkpoint"),
    "c2train": ("dss", 32, 0.075, "C2-train (SURVEY §8 row f-4): DSS baseline TRAINING step (30 unrolled layers with per-layer decoders and flux-residual losses; "
                                   "every layer = native fused forward + native layer backward psi_layer_backward), batch of 32 synthetic ~500-node meshes "
                                   "per GPU, shipped DSS checkpoint"),
    "c1": ("dirichlet", 32, 0.075, "C1: PSI-GNN dirichlet training step, batch of 32 synthetic ~500-node 2D triangle Poisson meshes"),
    "c3": ("dirichlet", 256, 0.075, "C3: PSI-GNN dirichlet training step (Broyden forward + implicit-adjoint backward), batch 256 synthetic ~500-node meshes per GPU"),
    "c5": ("dirichlet", 1, 0.075, "C5: PSI-GNN dirichlet forward Broyden solve (500-step cap) of ONE synthetic 1M-node mesh, node-range partitioned "
                                   "over the GPUs (halo rows and inner products exchanged through peer-mapped memory over NVLink)"),
    "c4": ("mixed", 256, 0.037, "C4: PSI-GNN mixed Dirichlet/Neumann training step, batch 256 synthetic ~2k-node meshes per GPU"),
}
C5_NODES = 1_000_000
# timing rule: at least 3 untimed warm-up steps whatever --warmup says (PSI_BENCH_MIN_WARMUP=1 only for short profiler runs)
MIN_WARMUP = int(os.environ.get("PSI_BENCH_MIN_WARMUP", "3"))
LR = 1e-6            # end-of-training learning rate (both arms)
CLIP = 0.1           # launch_local.sh --gradient_clip
JAC_WEIGHT = 1.0     # launch_local.sh --jac_weight


def load_params(family):
    z = np.load(os.path.join(ROOT, "tests", "golden", "%s_ckpt.npz" % family))
    P = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param.")}
    cfg = dict(latent_dim=10, hidden_dim=10, n_layers=1, fw_tol=float(z["cfg.fw_tol"]), fw_thres=int(z["cfg.fw_thres"]),
               bw_tol=float(z["cfg.bw_tol"]), bw_thres=int(z["cfg.bw_thres"]), path_logs=None)
    return P, cfg


def make_batch(family, n_graphs, h, seed0):
    from psi_gnn_b200 import synthetic
    return synthetic.make_batch(n_graphs, seed0=seed0, h=h, mixed=(family == "mixed"), solve=False)


def ncu_traffic(kernel, alg_bytes_per_launch):
    """ESTIMATE of dram__bytes_read+write per launch for the dominant kernel (reported as `traffic_estimate`, never as `traffic`): the ncu
    --set full capture in profiles/ncu_traffic.json gives the ratio of DRAM traffic to algorithmic bytes at ONE history depth; both scale
    with the depth, so the ratio is carried over to the average launch of this run.  It is not measured in this run and cannot reveal a
    traffic regression — re-capture with ncu after a kernel change (profiles/README.md)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return round(alg_bytes_per_launch * float(t[kernel]["ratio"]), 1)
    except Exception:
        return None


def training_config(desc, family, n_graphs, batch, cfg, world):
    """`config` of the training-step workloads — built identically by the native arm and by the reference arm (same workload; the
    reference arm's bounded sample is described in its `cpu_baseline.sample`)"""
    E = int((batch.edge_index[0] != batch.edge_index[1]).sum())
    return {"workload": desc, "graphs_per_gpu": n_graphs, "nodes_per_gpu": int(batch.num_nodes), "nnz_per_gpu": int(batch.edge_index.shape[1]),
            "offdiag_edges_per_gpu": E, "solver": "broyden", "fw_tol": cfg["fw_tol"], "fw_thres": cfg["fw_thres"], "bw_tol": cfg["bw_tol"],
            "bw_thres": cfg["bw_thres"], "jac_weight": JAC_WEIGHT, "lr": LR,
            "parallelism": "graph-sharded dp%d, one NCCL all-reduce of the flat gradient per step" % world,
            "l2": "256 MB buffer written between timed steps; the U/V history (GBs) exceeds the 126 MB L2 anyway"}


def inference_config(desc, mesh, cfg, world, one_mesh, n_owned, n_ghost):
    E_glob = int((mesh.edge_index[0] != mesh.edge_index[1]).sum())
    return {"workload": desc, "nodes": int(mesh.num_nodes), "nnz": int(mesh.edge_index.shape[1]), "offdiag_edges": E_glob,
            "owned_nodes_rank0": int(n_owned), "ghost_nodes_rank0": int(n_ghost), "solver": "broyden", "fw_tol": cfg["fw_tol"], "fw_thres": cfg["fw_thres"],
            "parallelism": ("node-range mesh partition x%d" % world if one_mesh else "independent batches x%d" % world) if world > 1 else "single GPU",
            "l2": "256 MB buffer written between timed steps; working set (history) exceeds L2"}


def dss_config(desc, n_graphs, batch, k):
    return {"workload": desc, "graphs_per_gpu": n_graphs, "nodes_per_gpu": int(batch.num_nodes), "edges_per_gpu": int(batch.edge_index.shape[1]),
            "layers": k, "l2": "256 MB buffer written between timed steps"}


# graphs of the bounded CPU sample of each training workload (first graphs of rank 0's batch: same seeds, same generator)
CPU_SAMPLE_GRAPHS = {"c1": 32, "c3": 32, "c4": 4}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# =====================================================================================================================
# native arm
# =====================================================================================================================
def run_native(args):
    import torch.distributed as dist
    from psi_gnn_b200 import _native, parallel
    from psi_gnn_b200 import model as PM

    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _native.load()                                              # fail loudly if the extension is missing
    family, n_graphs, h, desc = WORKLOADS[args.workload]
    if args.workload in ("c5", "c0"):
        return run_native_inference(args, rank, local, world, dev)
    if args.workload in ("c2", "c2train"):
        return run_native_dss(args, rank, local, world, dev)
    if args.graphs:
        n_graphs = args.graphs
    P, cfg = load_params(family)
    if family == "mixed":
        from psi_gnn_b200.mixed.psignn import model as M
        from psi_gnn_b200.mixed.psignn.utilities import solver as S
    else:
        from psi_gnn_b200.dirichlet.psignn import model as M
        from psi_gnn_b200.dirichlet.psignn.utilities import solver as S
    cfg["solver"] = S.broyden
    model = M.ModelDEQDSS(cfg)
    model.load_state_dict(P)
    model = model.to(dev).train()
    params = [p for p in model.parameters()]
    opt_deq = torch.optim.Adam(model.deqdss.parameters(), lr=LR)
    opt_ae = torch.optim.Adam(model.autoencoder.parameters(), lr=LR)

    host_batch = make_batch(family, n_graphs, h, seed0=rank * n_graphs).pin_memory()
    dev_batch = host_batch.to(dev)
    N, nnz = host_batch.num_nodes, host_batch.edge_index.shape[1]
    h2d = host_batch.nbytes()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)      # 256 MB > 126 MB L2

    stats = {"fw_steps": 0, "bw_steps": 0, "fw_evals": 0, "bw_evals": 0, "launches": 0}

    def train_step(batch, read_loss):
        opt_ae.zero_grad(set_to_none=True)
        opt_deq.zero_grad(set_to_none=True)
        _, ld = model(batch)
        loss = ld["residual_loss"].mean() + JAC_WEIGHT * ld["jacobian_loss"].mean() + ld["encoder_loss"].mean() + ld["autoencoder_loss"].mean()
        loss.backward()
        parallel.allreduce_gradients(params, world)
        torch.nn.utils.clip_grad_norm_(params, CLIP)
        opt_deq.step()
        opt_ae.step()
        fw, bw = model.deqdss.last_forward, model.deqdss.last_backward
        stats["fw_steps"] += fw["steps_run"]; stats["bw_steps"] += bw["steps_run"]
        stats["fw_evals"] += fw["f_evals"]; stats["bw_evals"] += bw["f_evals"]
        # native launches: both solves + vjp_prepare (1) + residual (2) + Aᵀr (1)
        stats["launches"] += fw["launches"] + bw["launches"] + 4
        return loss.item() if read_loss else None

    def e2e_step():
        b = host_batch.to(dev, non_blocking=True)               # pinned host → device copy of the step's inputs
        return train_step(b, True)                              # includes the graph re-layout and the D2H read of the loss

    def timed(fn, k):
        """K steps, each bracketed by CUDA events on the current stream, L2 flushed between steps (outside the brackets)"""
        evs = []
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for _ in range(k):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm --------------------------------------------------------------------------------
    warm = max(args.warmup, MIN_WARMUP)
    for _ in range(warm):
        train_step(dev_batch, False)
    ws = PM.graph_of(dev_batch, model.deqdss.f.kind).solver(max(cfg["fw_thres"], cfg["bw_thres"]))
    for k in stats:
        stats[k] = 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(lambda: train_step(dev_batch, False), args.steps)           # headline pass: no instrumentation inside the loop
    clocks = sampler.stop() if rank == 0 else None
    run_stats = dict(stats)
    # second pass of the same K steps with one CUDA-event pair around every solver-loop launch (psi_solver_profile) for the roofline:
    # the ~3 µs per event pair cost 2–3 % of a C3 step and 30 % of a C0 step, so they are kept out of the headline pass
    ws.profile(True)
    ms_profiled = timed(lambda: train_step(dev_batch, False), args.steps)
    prof = ws.profile_read()
    ws.profile(False)
    # ---- end-to-end arm (host buffers, H2D + re-layout + D2H inside the timed region) --------------------------
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)

    g = PM.graph_of(dev_batch, model.deqdss.f.kind)
    E = g.num_offdiag
    lists = 3 if family == "mixed" else 2
    sec = ms_total / 1e3
    # per-rank solver steps of the headline pass: the slowest rank sets the pace of a weak-scaling step (step time ∝ steps²)
    per_rank = torch.tensor([run_stats["fw_steps"], run_stats["bw_steps"], N], dtype=torch.float64, device=dev)
    gathered = [torch.zeros_like(per_rank) for _ in range(world)]
    if world > 1:
        dist.all_gather(gathered, per_rank)
    else:
        gathered = [per_rank]
    steps_per_rank = [{"rank": r, "forward": float(t[0]) / args.steps, "backward": float(t[1]) / args.steps, "nodes": int(t[2])}
                      for r, t in enumerate(gathered)]
    # ---- strong scaling of the SAME workload: the 256 graphs of rank 0's batch sharded over the ranks (32 graphs per GPU at 8) --
    strong = None
    if world > 1 and not args.no_extra_blocks:
        full = make_batch(family, n_graphs, h, seed0=0)
        shard = parallel.shard_batch(full, rank, world).to(dev)
        for _ in range(2):
            train_step(shard, False)
        for k in stats:
            stats[k] = 0
        ms_strong = timed(lambda: train_step(shard, False), args.steps)
        sst = torch.tensor([stats["fw_steps"], stats["bw_steps"], shard.num_graphs], dtype=torch.float64, device=dev)
        sg = [torch.zeros_like(sst) for _ in range(world)]
        dist.all_gather(sg, sst)
        strong = {"graphs_total": n_graphs, "value": round(n_graphs * args.steps / (ms_strong / 1e3), 2), "unit": "graphs/s",
                  "ms_per_step": round(ms_strong / args.steps, 3), "scaling": "strong",
                  "per_rank": [{"rank": r, "graphs": int(t[2]), "forward": float(t[0]) / args.steps, "backward": float(t[1]) / args.steps}
                               for r, t in enumerate(sg)],
                  "note": "same 256 graphs as the 1-GPU headline, contiguous node-balanced groups (parallel.shard_batch); compare with the "
                          "N=1 line's value for the strong-scaling speed-up"}
        del shard, full
    graphs_s = world * n_graphs * args.steps / sec
    iters = run_stats["fw_steps"] + run_stats["bw_steps"]
    evals = run_stats["fw_evals"] + run_stats["bw_evals"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    dom = max(prof, key=lambda k: prof[k]["ms"])
    d = prof[dom]
    achieved = d["bytes"] / max(d["ms"], 1e-9) / 1e6            # GB/s
    traffic = ncu_traffic(dom, d["bytes"] / max(d["launches"], 1))
    kernels = {k: {"launches": v["launches"], "ms_total": round(v["ms"], 3), "avg_us": round(1e3 * v["ms"] / max(v["launches"], 1), 2),
                   "alg_GBps": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1), "frac_of_peak": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6 / peak, 4)}
               for k, v in prof.items()}
    out = {
        "metric": "PSI-GNN solve graphs/s (training step: Broyden forward solve + implicit-adjoint backward solve)",
        "value": round(graphs_s, 2), "unit": "graphs/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic P1-FEM Poisson meshes (seeded generator); weights = reference shipped checkpoint",
        "config": training_config(desc, family, n_graphs, host_batch, cfg, world),
        "iterations_per_s": round(world * iters / sec, 1),
        "edge_msg_updates_per_s": round(world * evals * lists * E / sec, 1),
        "solver_steps_per_step": {"forward": run_stats["fw_steps"] / args.steps, "backward": run_stats["bw_steps"] / args.steps},
        "solver_steps_per_rank": steps_per_rank,
        "gpu_launches": run_stats["launches"],
        "e2e": {"value": round(world * n_graphs * args.steps / (ms_e2e / 1e3), 2), "unit": "graphs/s", "ms_per_step": round(ms_e2e / args.steps, 3),
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "measured_in": "second pass of the same K steps with one CUDA-event pair per solver-loop launch (%.3f ms/step instrumented vs %.3f ms/step in the headline pass)" % (ms_profiled / args.steps, ms_total / args.steps), "traffic": None, "traffic_estimate": traffic, "traffic_source": "NOT measured in this run: algorithmic bytes per launch x the DRAM-traffic/algorithmic ratio of the committed ncu --set full capture (profiles/ncu_traffic.json, one history depth)" if traffic else None, "peak_source": peak_src, "avg_launch_us": round(1e3 * d["ms"] / max(d["launches"], 1), 2),
                     "alg_bytes_per_launch": round(d["bytes"] / max(d["launches"], 1), 1)},
        "kernels": kernels,
        "clocks": clocks,
    }
    if strong is not None:
        out["strong"] = strong
    if args.workload == "c3" and not args.no_extra_blocks:
        # BASELINE configs[4] beside the headline: the 1M-node mesh solve, node-range partitioned over the same N GPUs
        del dev_batch, model
        _solver_release()
        try:
            out["mesh_partitioned"] = mesh_block(args, rank, local, world, dev)
        except Exception as exc:                                  # the headline line must survive a failure of the extra block
            out["mesh_partitioned"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args, family, h, budget_s=30.0)
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def _solver_release():
    from psi_gnn_b200 import solver as S
    S.release_workspaces()
    torch.cuda.empty_cache()


def mesh_block(args, rank, local, world, dev):
    """C5 (one 1M-node mesh, strong scaling over the N GPUs) as an extra block of the default line: forward solve (and, where the
    partitioned VJP is available, the implicit-adjoint backward solve) timed on the device, max over ranks"""
    sub = argparse.Namespace(**vars(args))
    sub.workload, sub.steps, sub.warmup, sub.no_cpu_baseline = "c5", max(2, min(args.steps, 3)), 1, True
    res = run_native_inference(sub, rank, local, world, dev, emit_line=False, min_warm=1)
    keep = ("value", "unit", "ms_per_step", "scaling", "solver_steps_per_step", "last_solve", "iterations_per_s", "edge_msg_updates_per_s",
            "kernels", "e2e", "backward", "comm")
    blk = {k: res[k] for k in keep if k in res}
    blk["config"] = {k: res["config"][k] for k in ("workload", "nodes", "nnz", "owned_nodes_rank0", "ghost_nodes_rank0", "parallelism")}
    blk["roofline"] = {k: res["roofline"][k] for k in ("kernel", "achieved", "frac", "avg_launch_us")}
    # the partitioned solve retraces the single-GPU trajectory step for step (aligned cuts + fp64 cross-chunk sums): these two
    # numbers must be identical at every N (compare the N=1 line's block)
    blk["retrace_key"] = {"steps_run": res["last_solve"]["steps_run"], "lowest": res["last_solve"]["lowest"]}
    return blk


def _timed_steps(fn, k, world, dev, flush):
    import torch.distributed as dist
    evs = []
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for _ in range(k):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_native_inference(args, rank, local, world, dev, emit_line=True, min_warm=None):
    """forward solves.  c5: one large mesh, unpartitioned on 1 GPU, node-range partitioned on N GPUs (strong scaling);
    c0: a batch of 32 meshes per GPU (weak scaling, independent solves per rank)."""
    import torch.distributed as dist
    from psi_gnn_b200 import model as PM, partition, synthetic
    from psi_gnn_b200.dirichlet.psignn import model as M
    from psi_gnn_b200.dirichlet.psignn.utilities import solver as S
    family, n_graphs, h, desc = WORKLOADS[args.workload]
    one_mesh = args.workload == "c5"
    nodes = args.nodes or C5_NODES
    P, cfg = load_params(family)
    cfg["solver"] = S.broyden
    model = M.ModelDEQDSS(cfg)
    model.load_state_dict(P)
    model = model.to(dev).eval()
    if one_mesh:
        mesh = synthetic.make_large_mesh(nodes, seed=0, h=h)
    else:
        n_graphs = args.graphs or n_graphs
        mesh = make_batch(family, n_graphs, h, seed0=rank * n_graphs)
    if world > 1 and one_mesh:
        part = partition.partition_mesh(mesh, world, rank=rank)[0]
        host = part.local.pin_memory()
        n_owned, n_ghost = part.n_owned, part.n_ghost
    elif one_mesh:
        # the same geometric node order the partitioner uses: with its 2048-node-aligned cuts the N-GPU solve then retraces this one step
        # for step (identical reductions), so the step count — and the work — is the same at every GPU count
        host = partition.reorder_mesh(mesh).pin_memory()
        n_owned, n_ghost = mesh.num_nodes, 0
    else:
        host = mesh.pin_memory()
        n_owned, n_ghost = mesh.num_nodes, 0
    dev_batch = host.to(dev)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    stats = {"steps": 0, "evals": 0, "launches": 0}

    def solve(batch, read):
        with torch.no_grad():
            u = model.inference(batch)
        out = model.deqdss.last_forward
        stats["steps"] += out["steps_run"]; stats["evals"] += out["f_evals"]; stats["launches"] += out["launches"] + 2
        return u.cpu() if read else None

    def e2e():
        b = host.to(dev, non_blocking=True)
        if world > 1:
            b.partition.comm = dev_batch.partition.comm        # the communicator is built once per process
        return solve(b, True)

    warm = max(args.warmup, MIN_WARMUP if min_warm is None else min_warm)
    for _ in range(warm):
        solve(dev_batch, False)
    g = PM.graph_of(dev_batch, 0)
    ws = g.solver(cfg["fw_thres"])
    for k in stats:
        stats[k] = 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = _timed_steps(lambda: solve(dev_batch, False), args.steps, world, dev, flush)     # headline pass: no instrumentation
    clocks = sampler.stop() if rank == 0 else None
    run_stats = dict(stats)
    ws.profile(True)                                                                            # second pass with per-launch CUDA events
    ms_profiled = _timed_steps(lambda: solve(dev_batch, False), args.steps, world, dev, flush)
    prof = ws.profile_read()
    ws.profile(False)
    e2e()
    ms_e2e = _timed_steps(e2e, args.steps, world, dev, flush)
    backward = None
    if one_mesh:
        # the implicit-adjoint solve y = Jᵀy + grad at the fixed point just found (native VJP kernels; on a partition the ghost rows of
        # S̄ are exchanged between the two VJP phases), with a deterministic cotangent that is the same function of the global node id at
        # every GPU count
        from psi_gnn_b200 import solver as SV
        gids = np.concatenate([host.partition.owned_global, host.partition.ghost_global]) if getattr(host, "partition", None) is not None \
            else np.arange(host.num_nodes)
        gid = torch.from_numpy(gids.astype(np.float32)).to(dev)
        grad = (1e-3 * torch.sin(0.37 * gid[:, None] + torch.arange(10, device=dev)[None].float())).contiguous()
        with torch.no_grad():
            model.inference(dev_batch)
        h_star = model.deqdss.last_forward["result"]
        op = SV.VjpOperator(model.deqdss.f, h_star, dev_batch, grad)
        y0 = torch.zeros_like(grad)
        bstats = {}

        def bw_solve():
            o = SV.broyden(op, y0, threshold=cfg["bw_thres"], eps=cfg["bw_tol"])
            bstats.update(steps_run=o["steps_run"], lowest=o["lowest"], nstep=o["nstep"], launches=o["launches"])

        bw_solve()
        ms_bw = _timed_steps(bw_solve, max(1, min(args.steps, 2)), world, dev, flush)
        backward = {"ms_per_solve": round(ms_bw / max(1, min(args.steps, 2)), 3), "steps_run": bstats["steps_run"], "lowest": bstats["lowest"],
                    "bw_tol": cfg["bw_tol"], "bw_thres": cfg["bw_thres"],
                    "us_per_step": round(1e3 * ms_bw / max(1, min(args.steps, 2)) / max(bstats["steps_run"], 1), 2),
                    "cotangent": "1e-3·sin(0.37·global node id + channel)"}
    sec = ms_total / 1e3
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    dom = max(prof, key=lambda k: prof[k]["ms"])
    d = prof[dom]
    achieved = d["bytes"] / max(d["ms"], 1e-9) / 1e6
    traffic = ncu_traffic(dom, d["bytes"] / max(d["launches"], 1))
    kernels = {k: {"launches": v["launches"], "ms_total": round(v["ms"], 3), "avg_us": round(1e3 * v["ms"] / max(v["launches"], 1), 2),
                   "alg_GBps": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1), "frac_of_peak": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6 / peak, 4)}
               for k, v in prof.items()}
    E_glob = int((mesh.edge_index[0] != mesh.edge_index[1]).sum())
    units = 1 if one_mesh else world * n_graphs            # graphs solved per step over all ranks
    mult = 1 if one_mesh else world                        # independent solves: iteration / edge rates add up over the ranks
    out = {
        "metric": "PSI-GNN solve graphs/s (forward Broyden solve%s)" % (" of one large mesh" if one_mesh else ", inference"),
        "value": round(units * args.steps / sec, 4), "unit": "graphs/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True, "scaling": "strong" if one_mesh else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic P1-FEM Poisson mesh (seeded generator); weights = reference shipped checkpoint",
        "config": inference_config(desc, mesh, cfg, world, one_mesh, n_owned, n_ghost),
        "iterations_per_s": round(mult * run_stats["steps"] / sec, 1),
        "edge_msg_updates_per_s": round(mult * run_stats["evals"] * 2 * E_glob / sec, 1),
        "solver_steps_per_step": {"forward": run_stats["steps"] / args.steps},
        "last_solve": {k: model.deqdss.last_forward[k] for k in ("lowest", "nstep", "steps_run", "stop_reason", "prot_break")},
        "gpu_launches": run_stats["launches"],
        "e2e": {"value": round(units * args.steps / (ms_e2e / 1e3), 4), "unit": "graphs/s", "ms_per_step": round(ms_e2e / args.steps, 3),
                "h2d_bytes_per_step": host.nbytes(), "d2h_bytes_per_step": 4 * n_owned},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "measured_in": "second pass of the same K steps with one CUDA-event pair per solver-loop launch (%.3f ms/step instrumented vs %.3f ms/step in the headline pass)" % (ms_profiled / args.steps, ms_total / args.steps), "traffic": None, "traffic_estimate": traffic, "traffic_source": "NOT measured in this run: algorithmic bytes per launch x the DRAM-traffic/algorithmic ratio of the committed ncu --set full capture (profiles/ncu_traffic.json, one history depth)" if traffic else None, "peak_source": peak_src, "avg_launch_us": round(1e3 * d["ms"] / max(d["launches"], 1), 2),
                     "alg_bytes_per_launch": round(d["bytes"] / max(d["launches"], 1), 1)},
        "kernels": kernels, "clocks": clocks,
    }
    if backward is not None:
        out["backward"] = backward
    if world > 1 and one_mesh:
        from psi_gnn_b200 import partition as PT
        out["comm"] = {"transport": "peer-mapped mailboxes (CUDA IPC over NVLink): halo rows stored into the consumer's memory, inner products "
                                    "exchanged inside the reduction kernel" if PT.USE_P2P else "grouped ncclSend/ncclRecv + ncclAllReduce",
                       "halo_rows_sent_rank0": int(len(host.partition.send_index)), "halo_rows_received_rank0": int(n_ghost),
                       "nvlink_bytes_per_step_rank0": int(len(host.partition.send_index)) * 48 + (world - 1) * 8 * (3 * 60 + 4),
                       "note": "per Broyden step: 48 B per sent halo row + the 3(n−1)+4 fp64 sums to each other rank (n ≈ 60 on average)"}
    if not emit_line:
        return out
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args, family, h, budget_s=25.0)
        emit(out)
    if world > 1:
        dist.destroy_process_group()


DSS_TRAIN_METRIC = "DSS baseline training-step graphs/s (30 unrolled layers with per-layer losses, forward + backward)"


def load_dss():
    z = np.load(os.path.join(ROOT, "tests", "golden", "dss_ckpt.npz"))
    P = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param.")}
    return P, dict(latent_dim=10, k=int(z["cfg.k"]), alpha=float(z["cfg.alpha"]), gamma=0.9)


def dss_batch(n_graphs, h, seed0):
    from psi_gnn_b200 import synthetic
    return synthetic.to_dss(synthetic.make_batch(n_graphs, seed0=seed0, h=h, solve=False))


def run_native_dss(args, rank, local, world, dev):
    """DSS baseline inference: k = 30 launches of the fused layer kernel (kind DSS) with per-layer weight blocks + one decode"""
    import torch.distributed as dist
    from psi_gnn_b200.dirichlet.dss import model as M
    _, n_graphs, h, desc = WORKLOADS["c2"]
    n_graphs = args.graphs or n_graphs
    P, cfg = load_dss()
    model = M.DeepStatisticalSolver(cfg)
    model.load_state_dict(P)
    train = args.workload == "c2train"
    if train:
        desc = WORKLOADS["c2train"][3]
    model = model.to(dev)
    model = model.train() if train else model.eval()
    host = dss_batch(n_graphs, h, rank * n_graphs).pin_memory()
    dev_batch = host.to(dev)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    N, nnz = host.num_nodes, int(host.edge_index.shape[1])
    warm = max(args.warmup, MIN_WARMUP)

    def train_step(batch):
        # forward (30 native layer launches + per-layer decoders and flux residuals) and backward (30 × psi_layer_backward)
        model.zero_grad(set_to_none=True)
        _, ld = model(batch)
        ld["train_loss"].backward()
        return ld["train_loss"].detach()

    step_fn = (lambda: train_step(dev_batch)) if train else (lambda: model.inference(dev_batch))
    for _ in range(warm):
        step_fn()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # per-launch time of the layer kernel: CUDA events around the 30 layer launches of each step (the decode is outside the bracket)
    ms_total = _timed_steps(step_fn, args.steps, world, dev, flush)
    clocks = sampler.stop() if rank == 0 else None

    def e2e():
        if train:
            return train_step(host.to(dev, non_blocking=True)).cpu()
        return model.inference(host.to(dev, non_blocking=True)).cpu()

    e2e()
    ms_e2e = _timed_steps(e2e, args.steps, world, dev, flush)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    k = cfg["k"]
    alg = N * (40.0 + 1.0 + 12.0 + 0.5 + 40.0) + 2 * nnz * 16.0          # h read, tag, b', offsets, h' written; 16-byte records per edge per list
    per_launch_us = 1e3 * ms_total / args.steps / k                         # upper bound: includes the weight upload and launch gaps
    achieved = alg / (per_launch_us * 1e-6) / 1e9
    sec = ms_total / 1e3
    out = {"metric": DSS_TRAIN_METRIC if train else "DSS baseline inference graphs/s (30 layers on the shared fused layer kernel)",
           "value": round(world * n_graphs * args.steps / sec, 2), "unit": "graphs/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
           "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic P1-FEM Poisson meshes in the DSS reader's layout; weights = reference shipped DSS checkpoint",
           "config": dss_config(desc, n_graphs, host, k),
           "edge_msg_updates_per_s": round(world * k * 2 * nnz * args.steps / sec, 1),
           # training: per layer pre-pass + layer (forward), 2 × k_bl_pass + k_bl_gather + k_pgrad_reduce (backward), k_flux forward and adjoint
           "gpu_launches": args.steps * (k * 8 + 2) if train else args.steps * (k + 1),
           "e2e": {"value": round(world * n_graphs * args.steps / (ms_e2e / 1e3), 2), "unit": "graphs/s", "ms_per_step": round(ms_e2e / args.steps, 3),
                   "h2d_bytes_per_step": host.nbytes(), "d2h_bytes_per_step": 4 if train else 4 * N},
           "roofline": {"bound": "hbm", "kernel": "k_layer_forward<DSS>", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                        "frac": round(achieved / peak, 4), "traffic": None, "avg_launch_us": round(per_launch_us, 2), "alg_bytes_per_launch": alg,
                        "note": ("training step: host-bound (30 unrolled layers x [layer kernels + torch decoder/loss block]); 'avg_launch_us' is the step time "
                                 "per unrolled layer, not a kernel duration" if train else
                                 "latency-bound at this size: 16 k nodes are a fraction of one wave; time per launch includes the 12.7 KB weight-block "
                                 "upload and the launch gap")},
           "clocks": clocks}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args, "dss", h, budget_s=15.0)
        emit(out)
    if world > 1:
        dist.destroy_process_group()


# =====================================================================================================================
# CPU arm: the oracle port of the reference's training step (the reference itself is Python that needs PyG/torch_sparse and
# /root/reference, neither of which exists on the GPU box)
# =====================================================================================================================
def cpu_step_fn(family, h, sample_graphs):
    from oracle import psignn_oracle as O
    P, cfg = load_params(family)
    batch = make_batch(family, sample_graphs, h, seed0=0)
    params = {k: v.clone() for k, v in P.items()}
    plist = [torch.nn.Parameter(v) for v in params.values()]
    opt = torch.optim.Adam(plist, lr=LR)
    gen = torch.Generator().manual_seed(0)

    def step():
        cur = {k: p.detach() for k, p in zip(params.keys(), plist)}
        v = torch.randn(batch.num_nodes, 10, generator=gen)
        out = O.training_forward_backward(cur, batch, cfg["fw_thres"], cfg["fw_tol"], cfg["bw_thres"], cfg["bw_tol"], v,
                                          jac_weight=JAC_WEIGHT, mixed=(family == "mixed"))
        for k, p in zip(params.keys(), plist):
            p.grad = out["grads"][k]
        torch.nn.utils.clip_grad_norm_(plist, CLIP)
        opt.step()
        fw_steps = len(out["fw"]["xest_trace"]) - 1
        bw_steps = len(out["bw"]["xest_trace"]) - 1
        return fw_steps, bw_steps

    return step, batch


def cpu_infer_fn(family, h, workload, sample):
    """oracle port of ModelDEQDSS.inference (encoder → Broyden → decoder) on a batch of `sample` meshes (c0) or one mesh of `sample` nodes (c5)"""
    from oracle import psignn_oracle as O
    from psi_gnn_b200 import synthetic
    P, cfg = load_params(family)
    batch = synthetic.make_large_mesh(sample, seed=0, h=h) if workload == "c5" else make_batch(family, sample, h, seed0=0)

    def step():
        out = O.inference(P, batch, cfg["fw_thres"], cfg["fw_tol"], mixed=(family == "mixed"))
        return len(out["xest_trace"]) - 1, 0

    return step, batch


def cpu_sample(args):
    """(step function, sample batch, graphs per step, description, config of the FULL workload) of the bounded CPU sample"""
    family, n_graphs, h, desc = WORKLOADS[args.workload]
    n_graphs = args.graphs or n_graphs
    if args.workload == "c2train":
        from oracle import psignn_oracle as O
        P, cfg = load_dss()
        P = {k_: v.clone().requires_grad_() for k_, v in P.items()}
        batch = dss_batch(n_graphs, h, 0)

        def step():
            # the reference's unrolled training forward + backward (dirichlet/dss/model.py:59-104, :129-148) as restated by the oracle
            for v in P.values():
                v.grad = None
            total, _, _ = O.dss_training_forward(P, batch, cfg["k"], cfg["alpha"], cfg["gamma"])
            total.backward()
            return cfg["k"], cfg["k"]

        return (step, batch, n_graphs, "one DSS training step (30 unrolled layers, forward + backward) of the full batch of %d meshes (N=%d nodes)"
                % (n_graphs, batch.num_nodes), dss_config(WORKLOADS["c2train"][3], n_graphs, batch, cfg["k"]))
    if args.workload == "c2":
        from oracle import psignn_oracle as O
        P, cfg = load_dss()
        batch = dss_batch(n_graphs, h, 0)

        def step():
            with torch.no_grad():
                O.dss_inference(P, batch, cfg["k"], cfg["alpha"])
            return cfg["k"], 0

        return (step, batch, n_graphs, "one DSS inference (30 layers) of the full batch of %d meshes (N=%d nodes)" % (n_graphs, batch.num_nodes),
                dss_config(desc, n_graphs, batch, cfg["k"]))
    _, cfg = load_params(family)
    if args.workload == "c5":
        from psi_gnn_b200 import synthetic
        step, batch = cpu_infer_fn(family, h, "c5", 20000)
        # the config of the full workload names the 1M-node mesh; its sizes come from the generator's closed form (no need to build it on the CPU arm)
        full = synthetic.make_large_mesh(args.nodes or C5_NODES, seed=0, h=h)
        return (step, batch, 1, "one forward solve of a %d-node mesh (the 1M-node mesh is out of reach of the CPU path in minutes)" % batch.num_nodes,
                inference_config(desc, full, cfg, 1, True, full.num_nodes, 0))
    if args.workload == "c0":
        step, batch = cpu_infer_fn(family, h, "c0", n_graphs)
        return (step, batch, n_graphs, "one forward solve of the full batch of %d meshes (N=%d nodes)" % (n_graphs, batch.num_nodes),
                inference_config(desc, batch, cfg, max(1, args.gpus), False, batch.num_nodes, 0))
    sample = min(n_graphs, CPU_SAMPLE_GRAPHS.get(args.workload, 8))
    step, batch = cpu_step_fn(family, h, sample)
    full = batch if sample == n_graphs else make_batch(family, n_graphs, h, seed0=0)
    what = ("one training step on the full batch of %d meshes (N=%d nodes)" % (sample, batch.num_nodes) if sample == n_graphs else
            "one training step on the first %d of the workload's %d meshes (N=%d of %d nodes; graphs/s on the sample flatters the CPU: the "
            "Broyden history cost per graph grows with the batch)" % (sample, n_graphs, batch.num_nodes, full.num_nodes))
    return step, batch, sample, what, training_config(desc, family, n_graphs, full, cfg, max(1, args.gpus))


def _time_steps(step, k):
    ts, its = [], 0
    for _ in range(k):
        t0 = time.perf_counter()
        fw, bw = step()
        ts.append(time.perf_counter() - t0)
        its += fw + bw
    return ts, its


def cpu_baseline(args, family, h, budget_s):
    """the oracle port on the host cores, on rank 0 at N=1: all cores, median of up to 5 steps inside the time budget"""
    torch.set_flush_denormal(True)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, batch, sample, what, _ = cpu_sample(args)
    step()                                                      # warm-up (library initialisation)
    t0 = time.perf_counter()
    ts, its = [], 0
    while len(ts) < 5 and (not ts or (time.perf_counter() - t0) + float(np.median(ts)) < budget_s):
        t, i = _time_steps(step, 1)
        ts += t
        its += i
    med = float(np.median(ts))
    return {"value": round(sample / med, 4), "unit": "graphs/s", "cores": torch.get_num_threads(), "kind": "port",
            "iterations_per_s": round(its / sum(ts), 2), "median_of": len(ts),
            "sample": "median of %d x %s; oracle port of the reference, torch %s CPU, %d threads; the 1-thread figure is in the "
                      "--impl reference line" % (len(ts), what, torch.__version__, torch.get_num_threads())}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    torch.set_flush_denormal(True)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, batch, sample, what, config = cpu_sample(args)
    warm = max(1, args.warmup)                                  # the same warm-up count as the native arm is asked for
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    ts, its = _time_steps(step, args.steps)
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    med = float(np.median(ts))
    # one step on a single thread (BASELINE.md §2 promises both figures)
    one = None
    if not args.no_cpu_baseline:
        torch.set_num_threads(1)
        t1, _ = _time_steps(step, 1)
        torch.set_num_threads(cores)
        one = round(sample / t1[0], 4)
    sample_txt = ("each step = %s; oracle port of the reference (the Python reference needs PyG/torch_sparse and cannot travel to the "
                  "GPU box); %d warm-up + %d timed steps, median step %.3f s" % (what, warm, args.steps, med))
    metric = {"c2": "DSS baseline inference graphs/s (30 layers on the shared fused layer kernel)", "c2train": DSS_TRAIN_METRIC,
              "c5": "PSI-GNN solve graphs/s (forward Broyden solve of one large mesh)", "c0": "PSI-GNN solve graphs/s (forward Broyden solve, inference)"}.get(
        args.workload, "PSI-GNN solve graphs/s (training step: Broyden forward solve + implicit-adjoint backward solve)")
    out = {"impl": "reference", "metric": metric,
           "value": round(val, 4), "unit": "graphs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
           "ms_per_step": round(1e3 * dt / args.steps, 2), "higher_is_better": True,
           "scaling": "strong" if args.workload == "c5" else "weak", "vs_baseline": None, "dtype": "f32",
           "data": ("synthetic P1-FEM Poisson mesh (seeded generator); weights = reference shipped checkpoint" if args.workload in ("c0", "c5") else
                    "synthetic P1-FEM Poisson meshes in the DSS reader's layout; weights = reference shipped DSS checkpoint" if args.workload in ("c2", "c2train") else
                    "synthetic P1-FEM Poisson meshes (seeded generator); weights = reference shipped checkpoint"),
           "config": config,
           "iterations_per_s": round(its / dt, 2),
           "cpu_baseline": {"value": round(val, 4), "unit": "graphs/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample_txt,
                            "sample_graphs_per_step": sample, "median_value": round(sample / med, 4), "median_of": len(ts),
                            "value_1thread": one},
           "e2e": {"value": round(val, 4), "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


_REAL_STDOUT = None


def emit(obj):
    """the ONE JSON line goes to the real stdout; everything else (NCCL banners, library chatter) was redirected to stderr"""
    line = json.dumps(obj) + "\n"
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, line.encode())
    else:
        sys.stdout.write(line)
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--graphs", type=int, default=0, help="override graphs per GPU")
    ap.add_argument("--nodes", type=int, default=0, help="override the mesh size of workload c5")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-blocks", action="store_true", help="skip the `strong` (C3, 256 graphs total) and `mesh_partitioned` (C5) blocks")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
