"""ctypes binding of the C ABI in ``include/psignn_b200.h``.

There is no fallback: if the shared library is missing or a call fails, a ``RuntimeError`` is
raised.  All pointers handed to the library are borrowed ``data_ptr()``s of torch CUDA tensors.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

import torch

from . import build as _build

_LIB = None


class SolveStats(ctypes.Structure):
    _fields_ = [("lowest", c_double), ("nstep", c_int32), ("steps_run", c_int32), ("prot_break", c_int32),
                ("stop_reason", c_int32), ("f_evals", c_int32), ("launches", c_int32)]


KIND_DIRICHLET, KIND_MIXED, KIND_DSS, KIND_DSGPS, KIND_DSGPS_MIXED = 0, 1, 2, 3, 4
OP_LAYER, OP_VJP = 0, 1

# name -> (restype, argtypes); every symbol include/psignn_b200.h declares
SIGNATURES = {
    "psi_version": (c_int, []),
    "psi_last_error": (c_char_p, []),
    "psi_weights_floats": (c_int, []),
    "psi_weights_upload": (c_int, [c_void_p, c_int, c_void_p]),
    "psi_graph_create": (c_int, [POINTER(c_void_p), c_int64, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                 c_void_p, c_int, c_void_p, c_void_p]),
    "psi_graph_destroy": (c_int, [c_void_p]),
    "psi_graph_info": (c_int, [c_void_p, POINTER(c_int64)]),
    "psi_comm_unique_id": (c_int, [ctypes.c_char_p]),
    "psi_comm_create": (c_int, [POINTER(c_void_p), c_int, c_int, ctypes.c_char_p]),
    "psi_comm_destroy": (c_int, [c_void_p]),
    "psi_graph_set_partition": (c_int, [c_void_p, c_void_p, c_int64, c_int, POINTER(c_int32), POINTER(c_int64), POINTER(c_int64), c_void_p, c_void_p]),
    "psi_halo_exchange": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "psi_part_mail_create": (c_int, [c_void_p, ctypes.c_char_p, POINTER(c_int64)]),
    "psi_part_mail_open": (c_int, [c_void_p, ctypes.c_char_p, POINTER(c_int64), POINTER(c_int64)]),
    "psi_part_error": (c_int, [c_void_p]),
    "psi_layer_forward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "psi_layers_unrolled": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "psi_vjp_prepare": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "psi_vjp_apply": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "psi_param_grad": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "psi_param_grad_tangent": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "psi_pgrad_layout": (c_int, [POINTER(c_int32)]),
    "psi_layer_backward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "psi_residual": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "psi_spmv_t": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "psi_flux": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "psi_encode": (c_int, [c_int64, c_void_p, c_void_p, c_void_p]),
    "psi_decode": (c_int, [c_int64, c_void_p, c_void_p, c_void_p]),
    "psi_solver_create": (c_int, [POINTER(c_void_p), c_int64, c_int]),
    "psi_solver_destroy": (c_int, [c_void_p]),
    "psi_solver_bytes": (c_int64, [c_void_p]),
    "psi_solver_stride": (c_int64, [c_void_p]),
    "psi_solver_profile": (c_int, [c_void_p, c_int]),
    "psi_solver_profile_read": (c_int, [c_void_p, POINTER(c_double)]),
    "psi_solver_broyden": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_double, c_void_p,
                                   POINTER(SolveStats), POINTER(c_double), POINTER(c_double), c_void_p, c_void_p]),
    "psi_broyden_begin": (c_int, [c_void_p, c_void_p, c_int, c_double, c_void_p, c_void_p]),
    "psi_broyden_x": (c_void_p, [c_void_p]),
    "psi_broyden_first": (c_int, [c_void_p, c_void_p, c_void_p]),
    "psi_broyden_step": (c_int, [c_void_p, c_void_p, POINTER(c_int), c_void_p]),
    "psi_broyden_finish": (c_int, [c_void_p, c_void_p, POINTER(SolveStats), POINTER(c_double), POINTER(c_double), c_void_p]),
    "psi_broyden_forced_step": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "psi_solver_anderson": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_double, c_int, c_double, c_double,
                                    c_void_p, POINTER(SolveStats), POINTER(c_double), POINTER(c_double), c_void_p, c_void_p]),
    "psi_solver_picard": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_double, c_void_p,
                                  POINTER(SolveStats), POINTER(c_double), POINTER(c_double), c_void_p, c_void_p]),
    "psi_anderson_begin": (c_int, [c_void_p, c_void_p, c_int, c_double, c_int, c_double, c_double, c_void_p, c_void_p]),
    "psi_anderson_x": (c_void_p, [c_void_p]),
    "psi_anderson_feed": (c_int, [c_void_p, c_void_p, POINTER(c_int), c_void_p]),
    "psi_anderson_finish": (c_int, [c_void_p, c_void_p, POINTER(SolveStats), POINTER(c_double), POINTER(c_double), c_void_p]),
    "psi_anderson_forced_step": (c_int, [c_void_p, c_int, c_int, c_int, c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "psi_picard_begin": (c_int, [c_void_p, c_void_p, c_int, c_double, c_void_p, c_void_p]),
    "psi_picard_x": (c_void_p, [c_void_p]),
    "psi_picard_feed": (c_int, [c_void_p, c_void_p, POINTER(c_int), c_void_p]),
    "psi_picard_finish": (c_int, [c_void_p, c_void_p, POINTER(SolveStats), POINTER(c_double), POINTER(c_double), c_void_p]),
}


def lib_path() -> str:
    return os.environ.get("PSI_GNN_B200_LIB", _build.LIB_PATH)


def load():
    """Load (once) the in-tree shared library and bind every declared symbol."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            "psi_gnn_b200: the CUDA extension %s is missing — run `python -m psi_gnn_b200.build` "
            "(there is no CPU or PyTorch fallback for the PSI-GNN solve)" % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().psi_last_error()
        raise RuntimeError("psi_gnn_b200 %s failed: %s" % (what, msg.decode() if msg else "unknown error"))


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """device pointer of a contiguous CUDA fp32/int64 tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("psi_gnn_b200: expected a CUDA tensor (the PSI-GNN hot path has no CPU implementation)")
    if not t.is_contiguous():
        raise RuntimeError("psi_gnn_b200: expected a contiguous tensor")
    return t.data_ptr()


def f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise RuntimeError("psi_gnn_b200: the native path computes in fp32 (got %s)" % t.dtype)
    return t.contiguous()
