"""CPU: the algebraic restructurings the CUDA kernels use (tests/restructured_math.py) change rounding order only."""
import torch

from conftest import Golden, rel_err
from restructured_math import f_dirichlet_restructured, vjp_dirichlet_restructured


def test_restructured_forward_fp32_and_fp64():
    for name in ("dirichlet_ckpt", "dirichlet_seed0"):
        g = Golden(name)
        P, b = g.params(), g.batch()
        h0 = g.t("h0")
        out = f_dirichlet_restructured(P, g.t("f1"), h0, b)
        assert rel_err(out, g.t("f2")) <= 1e-6
        P64 = {k: v.double() for k, v in P.items()}
        out64 = f_dirichlet_restructured(P64, g.t("f1").double(), h0.double(), b.double())
        assert rel_err(out64, g.t("f2")) <= 1e-6


def test_restructured_vjp():
    for name in ("dirichlet_ckpt", "dirichlet_seed0"):
        g = Golden(name)
        P, b = g.params(), g.batch()
        out = vjp_dirichlet_restructured(P, g.t("f2"), g.t("h0"), b, g.t("vjp_y"))
        assert rel_err(out, g.t("vjp_out")) <= 1e-5
