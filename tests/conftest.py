import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """One fixture of tests/golden (made by oracle/make_golden.py from the unmodified reference)."""

    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.mixed = "mixed" in name

    def __getitem__(self, k):
        return self.z[k]

    def t(self, k, device="cpu"):
        return torch.from_numpy(np.ascontiguousarray(self.z[k])).to(device)

    def has(self, k):
        return k in self.z.files

    def batch(self, device="cpu"):
        from psi_gnn_b200.synthetic import GraphData
        b = GraphData()
        for k in self.z.files:
            if k.startswith("batch.") and k != "batch.num_nodes":
                setattr(b, k[6:], self.t(k, device))
        b.num_nodes = int(self.z["batch.num_nodes"])
        b.num_graphs = int(b.ptr.numel() - 1) if getattr(b, "ptr", None) is not None else 1
        return b

    def params(self, device="cpu"):
        return {k[6:]: self.t(k, device) for k in self.z.files if k.startswith("param.")}

    def cfg(self):
        return dict(latent_dim=10, hidden_dim=10, n_layers=1, fw_tol=float(self.z["cfg.fw_tol"]), fw_thres=int(self.z["cfg.fw_thres"]),
                    bw_tol=float(self.z["cfg.bw_tol"]), bw_thres=int(self.z["cfg.bw_thres"]), path_logs=None)

    def model(self, device):
        """the drop-in ModelDEQDSS of this family with the fixture's weights"""
        if self.mixed:
            from psi_gnn_b200.mixed.psignn import model as M
            from psi_gnn_b200.mixed.psignn.utilities import solver as S
        else:
            from psi_gnn_b200.dirichlet.psignn import model as M
            from psi_gnn_b200.dirichlet.psignn.utilities import solver as S
        cfg = self.cfg()
        cfg["solver"] = S.broyden
        m = M.ModelDEQDSS(cfg)
        m.load_state_dict(self.params())
        return m.to(device)


ALL_FIXTURES = ["dirichlet_ckpt", "dirichlet_seed0", "mixed_ckpt", "mixed_seed0"]


@pytest.fixture(scope="session", params=ALL_FIXTURES)
def golden(request):
    return Golden(request.param)


@pytest.fixture(scope="session")
def golden_small():
    return Golden("dirichlet_ckpt_small")


def rel_err(a, b):
    a = a.double().cpu()
    b = b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))
