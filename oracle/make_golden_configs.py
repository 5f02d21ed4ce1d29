"""TEST INFRASTRUCTURE ONLY — golden vectors of the UNMODIFIED reference at the sizes of the BASELINE configurations.

Run in the build container (needs /root/reference):  ``python -m oracle.make_golden_configs [c1] [c3shard] [c4mixed]``

  cfg_c1        C0/C1: 32 ≈500-node meshes (seeds 0..31), shipped Dirichlet checkpoint — forward solve + one training step
  cfg_c3shard   one 8-GPU shard of C3: 32 meshes (seeds 224..255 = the last rank's share of the 256), forward solve only
  cfg_c4mixed   C4 sample: 8 ≈2 k-node mixed Dirichlet/Neumann meshes, shipped mixed checkpoint — forward solve + training step,
                including the reference's own stop reasons (does it run into the 500-step cap?)

Besides the reference's fp32 outputs (same keys as oracle/make_golden.py) every fixture stores what SURVEY §8c (3)/(4) asks for:
  perm_*        the reference run against itself with the edge list permuted (its own fp32 scatter)
  u64, hstar64  the fixed point solved by the reference's own code in fp64 to rel 1e-11 ("fp64-tight truth")
  train64_*     the training step in fp64 with both solves tightened (gradient truth), same Hutchinson probe v
  tf64_*        the same fp64 backward + gradients at the reference's fp32 H* (teacher-forced gradient truth)
"""
from __future__ import annotations

import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim                     # noqa: E402
from oracle.make_golden import BATCH_FIELDS, OUT, build_model   # noqa: E402
from psi_gnn_b200 import synthetic              # noqa: E402


def _loss(loss_dic):
    return loss_dic["residual_loss"].mean() + 1.0 * loss_dic["jacobian_loss"].mean() + loss_dic["encoder_loss"].mean() + \
        loss_dic["autoencoder_loss"].mean()


def _train_step(model, solver_mod, batch, v, fw=None, bw=None, pin_hstar=None):
    """ModelDEQDSS.forward + loss.backward() of the reference; records both solver calls.  ``fw``/``bw`` override (threshold, eps);
    ``pin_hstar`` replaces the forward solve's result (teacher forcing)."""
    rec = {"calls": []}
    orig = solver_mod.broyden

    def solver(fn, x0, threshold, eps):
        first = not rec["calls"]
        thr, tol = (fw if first else bw) or (threshold, eps)
        if first and pin_hstar is not None:
            out_ = {"result": pin_hstar.clone(), "lowest": 0.0, "nstep": 0, "xest_trace": [None], "rel_trace": []}
        else:
            out_ = orig(fn, x0, threshold=thr, eps=tol)
        rec["calls"].append(out_)
        if len(rec["calls"]) == 2:
            rec["bw_grad"] = fn(torch.zeros_like(x0)).detach().clone()      # y0 = 0 ⇒ fn(0) = grad (model.py:214-218)
        return out_

    model.config_deq["solver"] = solver
    model.deqdss.config_deq["solver"] = solver
    torch.cuda.synchronize = lambda *a, **k: None                            # model.py:213 on a CUDA-less build
    real_randn = torch.randn
    torch.randn = lambda *a, **k: v.clone()                                   # the probe of jac_loss_estimate (model.py:431)
    try:
        model.train()
        model.zero_grad()
        u, loss_dic = model(batch)
        _loss(loss_dic).backward()
    finally:
        torch.randn = real_randn
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)).detach().clone() for k, p in model.named_parameters()}
    return u.detach(), loss_dic, grads, rec


def make(family: str, name: str, n_graphs: int, seed0: int, h: float, train: bool, truth_thres: int = 1200):
    mixed = family.startswith("mixed")
    t0 = time.time()
    batch = synthetic.make_batch(n_graphs, seed0=seed0, h=h, mixed=mixed)
    fx = {}
    for k in BATCH_FIELDS:
        v = getattr(batch, k, None)
        if v is not None:
            fx["batch." + k] = v.numpy()
    fx["batch.num_nodes"] = np.int64(batch.num_nodes)
    torch.set_flush_denormal(True)
    with tempfile.TemporaryDirectory() as logdir:
        model, solver_mod, cfg = build_model(family, "ckpt", logdir)
        for k, v in model.state_dict().items():
            fx["param." + k] = v.detach().numpy().copy()
        for k in ("fw_tol", "fw_thres", "bw_tol", "bw_thres"):
            fx["cfg." + k] = np.float64(cfg[k])
        f = model.deqdss.f
        # ---- fp32 reference: forward solve ---------------------------------------------------------------------------
        with torch.no_grad():
            h0 = model.autoencoder.encoder(batch.x)
            fw = solver_mod.broyden(lambda Hh: f(Hh, h0, batch), h0, threshold=cfg["fw_thres"], eps=cfg["fw_tol"])
            u = model.autoencoder.decoder(fw["result"])
            res = model.residual_loss(u, batch)
        steps_run = len(fw["xest_trace"]) - 1
        fx.update(h0=h0.numpy(), fw_result=fw["result"].numpy(), fw_lowest=np.float64(fw["lowest"]), fw_nstep=np.int64(fw["nstep"]),
                  fw_steps_run=np.int64(steps_run), fw_prot_break=np.int64(bool(fw["prot_break"])),
                  fw_rel_trace=np.asarray(fw["rel_trace"], np.float64), u=u.numpy(), residual=np.float64(res.item()))
        print("  [%s] fp32 forward: nstep %d, steps run %d, lowest %.3e  (%.0f s)" % (name, fw["nstep"], steps_run, fw["lowest"], time.time() - t0))
        # ---- the reference against itself under an edge permutation ------------------------------------------------------
        perm = torch.randperm(batch.edge_index.shape[1], generator=torch.Generator().manual_seed(77))
        bp = synthetic.GraphData()
        bp.__dict__.update(batch.__dict__)
        bp.edge_index, bp.edge_attr, bp.a_ij = batch.edge_index[:, perm], batch.edge_attr[perm], batch.a_ij[perm]
        with torch.no_grad():
            fwp = solver_mod.broyden(lambda Hh: f(Hh, h0, bp), h0, threshold=cfg["fw_thres"], eps=cfg["fw_tol"])
            up = model.autoencoder.decoder(fwp["result"])
        fx.update(perm_u=up.numpy(), perm_fw_nstep=np.int64(fwp["nstep"]), perm_fw_lowest=np.float64(fwp["lowest"]),
                  perm_fw_steps_run=np.int64(len(fwp["xest_trace"]) - 1), perm_fw_rel_trace=np.asarray(fwp["rel_trace"], np.float64))
        print("  [%s] permuted reference: nstep %d, u rel diff to itself %.2e" % (name, fwp["nstep"], float((up - u).norm() / u.norm())))
        # ---- training step (fp32) ----------------------------------------------------------------------------------------
        v = torch.randn(h0.shape, generator=torch.Generator().manual_seed(4321))
        if train:
            u_tr, loss_dic, grads, rec = _train_step(model, solver_mod, batch, v)
            bw = rec["calls"][1]
            fx["train_v"] = v.numpy()
            fx["train_u"] = u_tr.numpy()
            for k, t in loss_dic.items():
                fx["train_loss." + k] = np.float64(t.item())
            for k, g in grads.items():
                fx["train_grad." + k] = g.numpy()
            fx.update(train_hstar=rec["calls"][0]["result"].detach().numpy(), train_bw_grad=rec["bw_grad"].numpy(),
                      train_fw_nstep=np.int64(rec["calls"][0]["nstep"]), train_fw_steps_run=np.int64(len(rec["calls"][0]["xest_trace"]) - 1),
                      train_bw_nstep=np.int64(bw["nstep"]), train_bw_steps_run=np.int64(len(bw["xest_trace"]) - 1),
                      train_bw_lowest=np.float64(bw["lowest"]), train_bw_result=bw["result"].numpy(),
                      train_bw_prot_break=np.int64(bool(bw["prot_break"])))
            print("  [%s] fp32 training step: fw %d steps, bw %d steps (best at %d, lowest %.2e)  (%.0f s)" % (
                name, fx["train_fw_steps_run"], fx["train_bw_steps_run"], bw["nstep"], bw["lowest"], time.time() - t0))
    # ---- fp64 truths (the reference's own code, default dtype switched so that broyden allocates its history in fp64) -----------
    torch.set_default_dtype(torch.float64)
    try:
        with tempfile.TemporaryDirectory() as logdir:
            model64, solver_mod, cfg = build_model(family, "ckpt", logdir)
            model64 = model64.double()
            b64 = batch.double()
            f64 = model64.deqdss.f
            with torch.no_grad():
                h064 = model64.autoencoder.encoder(b64.x)
                fw64 = solver_mod.broyden(lambda Hh: f64(Hh, h064, b64), h064, threshold=truth_thres, eps=1e-11)
                u64 = model64.autoencoder.decoder(fw64["result"])
            fx.update(hstar64=fw64["result"].numpy(), u64=u64.numpy(), fw64_lowest=np.float64(fw64["lowest"]),
                      fw64_nstep=np.int64(fw64["nstep"]))
            d32 = float((u.double() - u64).norm() / u64.norm())
            dp = float((up.double() - u64).norm() / u64.norm())
            print("  [%s] fp64-tight fixed point: %d steps, lowest %.2e | reference fp32 vs truth %.3e, permuted reference vs truth %.3e  (%.0f s)" % (
                name, fw64["nstep"], fw64["lowest"], d32, dp, time.time() - t0))
            if train:
                v64 = v.double()
                # (a) everything in fp64, both solves tight: the gradient truth of a free-running step
                _, ld64, g64, rec64 = _train_step(model64, solver_mod, b64, v64, fw=(truth_thres, 1e-11), bw=(truth_thres, 1e-12))
                for k, g in g64.items():
                    fx["train64_grad." + k] = g.numpy()
                for k, t in ld64.items():
                    fx["train64_loss." + k] = np.float64(t.item())
                fx["train64_bw_lowest"] = np.float64(rec64["calls"][1]["lowest"])
                # (b) fp64 backward + gradients at the reference's fp32 H*: the gradient truth of the teacher-forced step
                hs = torch.from_numpy(fx["train_hstar"]).double()
                u_tf, ldtf, gtf, rectf = _train_step(model64, solver_mod, b64, v64, bw=(truth_thres, 1e-12), pin_hstar=hs)
                for k, g in gtf.items():
                    fx["tf64_grad." + k] = g.numpy()
                for k, t in ldtf.items():
                    fx["tf64_loss." + k] = np.float64(t.item())
                fx["tf64_bw_result"] = rectf["calls"][1]["result"].numpy()
                fx["tf64_bw_lowest"] = np.float64(rectf["calls"][1]["lowest"])
                cat = lambda d_, pre: torch.cat([torch.from_numpy(np.asarray(d_[pre + k])).reshape(-1).double() for k in g64])
                g32 = cat(fx, "train_grad.")
                print("  [%s] gradient: reference fp32 vs fp64 truth %.3e (free-running), vs teacher-forced truth %.3e ; bw lowest %.1e / %.1e  (%.0f s)" % (
                    name, float((g32 - cat(fx, "train64_grad.")).norm() / cat(fx, "train64_grad.").norm()),
                    float((g32 - cat(fx, "tf64_grad.")).norm() / cat(fx, "tf64_grad.").norm()),
                    fx["train64_bw_lowest"], fx["tf64_bw_lowest"], time.time() - t0))
    finally:
        torch.set_default_dtype(torch.float32)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **fx)
    print("%-14s N=%d nnz=%d -> %s (%.0f kB, %.0f s)" % (name, batch.num_nodes, batch.edge_index.shape[1], os.path.relpath(path, ROOT),
                                                        os.path.getsize(path) / 1024, time.time() - t0))


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    which = sys.argv[1:] or ["c1", "c3shard", "c4mixed"]
    if "c1" in which:
        make("dirichlet/psignn", "cfg_c1", 32, 0, 0.075, train=True)
    if "c3shard" in which:
        make("dirichlet/psignn", "cfg_c3shard", 32, 224, 0.075, train=False)
    if "c4mixed" in which:
        make("mixed/psignn", "cfg_c4mixed", 8, 0, 0.037, train=True, truth_thres=2000)


if __name__ == "__main__":
    main()
