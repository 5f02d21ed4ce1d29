"""Parity numbers of the CUDA path against every golden fixture, in the layers of SURVEY §8c (run on the GPU box):

    python scripts/parity_report.py > profiles/r02_parity_report.txt

For each fixture: single applications, the free-running forward solve (first rel-trace entries, step counts, deviation of u from the
reference, from the permuted reference and — where the fixture holds it — from the fp64-tight fixed point), and the training-step
gradients against the reference / the fp64 truths.  The assertions live in tests/test_gpu_parity.py; this prints the raw numbers."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DEV = "cuda:0"


def main():
    from conftest import Golden, rel_err
    from psi_gnn_b200 import model as PM
    from psi_gnn_b200 import solver as S
    names = ["dirichlet_ckpt", "dirichlet_seed0", "mixed_ckpt", "mixed_seed0", "cfg_c1", "cfg_c3shard", "cfg_c4mixed"]
    for name in names:
        if not os.path.exists(os.path.join(ROOT, "tests", "golden", name + ".npz")):
            continue
        g = Golden(name)
        m = g.model(DEV)
        b = g.batch(DEV)
        h0 = g.t("h0", DEV)
        out = m.deqdss.inference(h0, b)
        u = m._decode_native(out["result"])
        ref_rel = g["fw_rel_trace"]
        k = min(30, int(g["fw_steps_run"]), out["steps_run"])
        dev_tr = np.abs(np.asarray(out["rel_trace"][:k]) - ref_rel[:k]) / ref_rel[:k]
        first_bad = int(np.argmax(dev_tr > 1e-3)) if (dev_tr > 1e-3).any() else k
        line = "%-16s N=%6d | nstep ours %3d ref %3d perm %3d | lowest %.2e (ref %.2e) | rel-trace within 1e-3 for the first %d steps" % (
            name, b.num_nodes, out["nstep"], int(g["fw_nstep"]), int(g["perm_fw_nstep"]), out["lowest"], float(g["fw_lowest"]), first_bad)
        if g.has("perm_fw_rel_trace"):
            pr = g["perm_fw_rel_trace"]
            kk = min(30, len(pr), len(ref_rel))
            d = np.abs(pr[:kk] - ref_rel[:kk]) / ref_rel[:kk]
            line += " (reference vs permuted reference: %d)" % (int(np.argmax(d > 1e-3)) if (d > 1e-3).any() else kk)
        print(line)
        print("    u: ours vs ref %.3e | ref vs permuted ref %.3e" % (rel_err(u, g.t("u")), rel_err(g.t("perm_u"), g.t("u"))), end="")
        if g.has("u64"):
            print(" | vs fp64-tight truth: ours %.3e, ref %.3e, permuted ref %.3e" % (
                rel_err(u, g.t("u64")), rel_err(g.t("u"), g.t("u64")), rel_err(g.t("perm_u"), g.t("u64"))), end="")
        print()
        if g.has("u64"):
            # our own run-to-run spread: the same solve with the edge list permuted (other summation orders inside the SELL rows)
            from psi_gnn_b200.synthetic import GraphData
            devs, nst = [], []
            for k in range(8):
                perm = torch.randperm(b.edge_index.shape[1], generator=torch.Generator().manual_seed(2000 + k)).to(DEV)
                bp = GraphData()
                bp.__dict__.update({k_: v_ for k_, v_ in b.__dict__.items() if not k_.startswith("_psi")})
                bp.edge_index, bp.edge_attr, bp.a_ij = b.edge_index[:, perm].contiguous(), b.edge_attr[perm].contiguous(), b.a_ij[perm].contiguous()
                o2 = m.deqdss.inference(h0, bp)
                devs.append(rel_err(m._decode_native(o2["result"]), g.t("u64")))
                nst.append(o2["nstep"])
            print("    ours under 8 edge permutations: nstep %s | u vs fp64-tight truth %s" % (sorted(nst), " ".join("%.2e" % d for d in sorted(devs))))
            if g.has("spread_u_dev64"):
                print("    reference under 8 edge permutations: nstep %s | u vs truth %s" % (
                    sorted(int(x) for x in g["spread_nstep"]), " ".join("%.2e" % d for d in sorted(g["spread_u_dev64"]))))
        if not g.has("train_v"):
            continue
        # training step, free-running and teacher-forced
        real_randn = torch.randn
        for mode in ("free", "forced"):
            m = g.model(DEV)
            if mode == "forced":
                hstar = g.t("train_hstar", DEV)
                calls = []

                def solver(f, x0, threshold, eps, _calls=calls, _h=hstar):
                    if not _calls:
                        _calls.append(1)
                        return {"result": _h.clone(), "lowest": 0.0, "nstep": 0, "steps_run": 0, "f_evals": 0, "launches": 0}
                    return S.broyden(f, x0, threshold=threshold, eps=eps)

                m.deqdss.config_deq["solver"] = solver
            v = g.t("train_v", DEV)
            PM.torch.randn = lambda *a, **k: v.clone()
            try:
                m.train()
                m.zero_grad()
                _, ld = m(b)
                (ld["residual_loss"].mean() + ld["jacobian_loss"].mean() + ld["encoder_loss"].mean() + ld["autoencoder_loss"].mean()).backward()
            finally:
                PM.torch.randn = real_randn
            names_ = [k_ for k_, _ in m.named_parameters()]
            gs = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).double().cpu() for _, p in m.named_parameters()])
            cat = lambda pre: torch.cat([g.t(pre + k_).reshape(-1).double() for k_ in names_])
            r32 = cat("train_grad.")
            bw = m.deqdss.last_backward
            s = "    train[%s]: bw steps %d lowest %.2e (ref %d, %.2e) | grad vs ref32 %.3e cos %.6f" % (
                mode, bw["steps_run"], bw["lowest"], int(g["train_bw_steps_run"]) if g.has("train_bw_steps_run") else -1,
                float(g["train_bw_lowest"]), float((gs - r32).norm() / r32.norm()), float(gs @ r32 / (gs.norm() * r32.norm())))
            key = "train64_grad." if mode == "free" else "tf64_grad."
            if g.has(key + names_[0]):
                t64 = cat(key)
                s += " | vs fp64 truth: ours %.3e, ref32 %.3e" % (float((gs - t64).norm() / t64.norm()), float((r32 - t64).norm() / t64.norm()))
            print(s)
            for k_ in ("residual_loss", "jacobian_loss", "encoder_loss", "autoencoder_loss"):
                ref = float(g["train_loss." + k_])
                print("        %-17s ours %.6e ref %.6e (%.1e)" % (k_, ld[k_].item(), ref, abs(ld[k_].item() - ref) / abs(ref)))


if __name__ == "__main__":
    main()
