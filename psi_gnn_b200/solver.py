"""Fixed-point solvers with the call signatures of the reference's ``utilities/solver.py``.

``broyden(f, x0, threshold, eps=1e-3, stop_mode="rel", ls=False, name="unknown")`` (solver.py:116-207),
``anderson(f, x0, m=2, lam=1e-4, threshold=50, eps=1e-3, stop_mode='rel', beta=1.0)`` (:215-293) and
``forward_iteration(f, z0, eps=1e-5, threshold=50)`` (:301-341) return the reference's result dict.

``f`` may be
  * a :class:`LayerOperator` / :class:`VjpOperator` — what ``DeepEquilibrium`` passes.  The whole solve then runs
    as a device-resident loop of fused CUDA kernels (one operator kernel + four quasi-Newton kernels per
    step, no torch op and no host synchronisation inside the loop except a stop-flag poll every few steps);
  * any other callable on CUDA tensors — the same kernels are driven through the step APIs of the extension
    (``psi_broyden_*``, ``psi_anderson_*``, ``psi_picard_*``), with one Python call of ``f`` per step; no solver
    algebra runs as torch ops.
There is no CPU implementation: CPU tensors raise ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, byref, c_double, c_int, c_void_p
from typing import Callable, List, Optional

import torch

from . import _native as N

# The reference keeps every iterate alive in ``xest_trace`` ((threshold+1)·N·d floats).  Off by default here;
# set to True (or pass keep_trace=True) for ``iterative_inference``-style consumers.
KEEP_TRACE = False


class SolverWorkspace:
    """Owns a ``psi_solver_t``: iterate, residual, δx/δg, best iterate and the U/V history (lazily grown)."""

    def __init__(self, numel: int, max_threshold: int, device):
        self.numel, self.cap, self.device = int(numel), int(max_threshold), torch.device(device)
        self.handle = c_void_p()
        with torch.cuda.device(self.device):
            N.check(N.load().psi_solver_create(byref(self.handle), self.numel, self.cap), "psi_solver_create")
        self.stride = int(N.load().psi_solver_stride(self.handle))

    def close(self):
        if self.handle:
            N.load().psi_solver_destroy(self.handle)
            self.handle = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    PROFILE_CLASSES = ("operator", "k_qn_dots", "k_qn_axpy")

    def profile(self, enable: bool = True):
        """per-kernel-class CUDA-event timing of the solver loops (resets the totals when enabled)"""
        N.check(N.load().psi_solver_profile(self.handle, 1 if enable else 0), "psi_solver_profile")

    def profile_read(self) -> dict:
        buf = (c_double * 9)()
        N.check(N.load().psi_solver_profile_read(self.handle, buf), "psi_solver_profile_read")
        return {name: {"launches": int(buf[3 * i]), "ms": float(buf[3 * i + 1]), "bytes": float(buf[3 * i + 2])}
                for i, name in enumerate(self.PROFILE_CLASSES)}

    @property
    def nbytes(self) -> int:
        return int(N.load().psi_solver_bytes(self.handle))


_POOL = {}           # (device index, numel) -> SolverWorkspace ; bounded by POOL_MAX_BYTES (least recently used evicted)
_POOL_ORDER = []
POOL_MAX_BYTES = 64 << 30


def workspace_for(numel: int, threshold: int, device) -> SolverWorkspace:
    device = torch.device(device)
    key = (device.index if device.index is not None else torch.cuda.current_device(), int(numel))
    ws = _POOL.get(key)
    if ws is None or ws.cap < threshold:
        if ws is not None:
            ws.close()
        ws = SolverWorkspace(numel, max(int(threshold), 1), device)
        _POOL[key] = ws
    if key in _POOL_ORDER:
        _POOL_ORDER.remove(key)
    _POOL_ORDER.append(key)
    while len(_POOL_ORDER) > 1 and len(_POOL) > 1 and sum(w.nbytes for w in _POOL.values()) > POOL_MAX_BYTES:
        old = _POOL_ORDER.pop(0)
        _POOL.pop(old).close()
    return ws


def release_workspaces():
    for ws in _POOL.values():
        ws.close()
    _POOL.clear()
    _POOL_ORDER.clear()


class NativeOperator:
    """Base class of the operators whose solves run entirely inside the extension."""
    graph = None
    kind = 0
    op = N.OP_LAYER
    aux: Optional[torch.Tensor] = None

    def upload(self):
        pass


class LayerOperator(NativeOperator):
    """``lambda H: f(H, H_init, batch)`` of dirichlet/psignn/model.py:189,247 as an object the solvers recognise."""
    op = N.OP_LAYER

    def __init__(self, function, h_init: torch.Tensor, batch):
        self.function, self.batch = function, batch
        self.graph = function.native_graph(batch)
        self.kind = function.kind
        self.aux = N.f32(h_init.detach())

    def upload(self):
        self.function.upload_weights(self.aux.device)

    def __call__(self, h: torch.Tensor) -> torch.Tensor:
        self.upload()
        return self.graph.layer_forward(self.kind, h, self.aux)


class VjpOperator(NativeOperator):
    """``lambda y: autograd.grad(f(H*), H*, y)[0] + grad`` of model.py:214 at the frozen point H*."""
    op = N.OP_VJP

    def __init__(self, function, h_star: torch.Tensor, batch, grad: torch.Tensor):
        self.function, self.batch = function, batch
        self.graph = function.native_graph(batch)
        self.kind = function.kind
        self.aux = N.f32(grad.detach())
        self.upload()
        self.graph.vjp_prepare(self.kind, h_star.detach())

    def upload(self):
        self.function.upload_weights(self.aux.device)

    def __call__(self, y: torch.Tensor) -> torch.Tensor:
        self.upload()
        return self.graph.vjp_apply(self.kind, y, self.aux)


def _result_dict(result, stats: N.SolveStats, rel, abs_, trace, eps, threshold, extra=None):
    out = {"result": result, "lowest": float(stats.lowest), "nstep": int(stats.nstep), "prot_break": bool(stats.prot_break),
           "abs_trace": list(abs_), "rel_trace": list(rel), "xest_trace": trace, "eps": eps, "threshold": threshold,
           # additions (not in the reference dict)
           "steps_run": int(stats.steps_run), "f_evals": int(stats.f_evals), "launches": int(stats.launches),
           "stop_reason": int(stats.stop_reason)}
    if extra:
        out.update(extra)
    return out


def _require_cuda(x0: torch.Tensor):
    if not x0.is_cuda:
        raise RuntimeError("psi_gnn_b200.solver: CUDA tensors required — the solvers have no CPU implementation")
    if x0.dtype != torch.float32:
        raise RuntimeError("psi_gnn_b200.solver: fp32 only (the reference allocates its history in fp32, solver.py:134-135)")


def _trace_views(buf: Optional[torch.Tensor], count: int, numel: int, shape) -> List[torch.Tensor]:
    if buf is None:
        return []
    return [buf[i, :numel].view(shape) for i in range(count)]


def broyden(f: Callable, x0: torch.Tensor, threshold: int, eps: float = 1e-3, stop_mode: str = "rel", ls: bool = False,
            name: str = "unknown", keep_trace: Optional[bool] = None) -> dict:
    """Good-Broyden on g(x) = f(x) − x, whole batch as one vector, step length 1 (reference solver.py:116-207)."""
    if stop_mode != "rel":
        raise NotImplementedError("psi_gnn_b200.solver.broyden: only stop_mode='rel' (the mode every reference call site uses)")
    if ls:
        raise NotImplementedError("psi_gnn_b200.solver.broyden: line search is never enabled by the reference (ls=False)")
    _require_cuda(x0)
    lib = N.load()
    keep = KEEP_TRACE if keep_trace is None else keep_trace
    x0c = N.f32(x0.detach())
    shape, numel = x0c.shape, x0c.numel()
    threshold = int(threshold)
    stats = N.SolveStats()
    rel = (c_double * (threshold + 1))()
    abs_ = (c_double * (threshold + 1))()
    result = torch.empty_like(x0c)
    with torch.cuda.device(x0c.device):
        stream = N.stream_ptr()
        if isinstance(f, NativeOperator):
            ws = f.graph.solver(threshold)
            xtrace = torch.empty(threshold + 1, ws.stride, dtype=torch.float32, device=x0c.device) if keep else None
            f.upload()
            N.check(lib.psi_solver_broyden(ws.handle, f.graph.handle, f.kind, f.op, N.ptr(x0c), N.ptr(f.aux), threshold, float(eps),
                                           N.ptr(result), byref(stats), rel, abs_, N.ptr(xtrace), stream), "psi_solver_broyden")
        else:
            ws = SolverWorkspace(numel, max(threshold, 1), x0c.device)
            xtrace = torch.empty(threshold + 1, ws.stride, dtype=torch.float32, device=x0c.device) if keep else None
            try:
                N.check(lib.psi_broyden_begin(ws.handle, N.ptr(x0c), threshold, float(eps), N.ptr(xtrace), stream), "psi_broyden_begin")
                # the iterate lives in the workspace; expose it to f as a tensor view without copying
                xview = _as_tensor(lib.psi_broyden_x(ws.handle), numel, x0c.device).view(shape)
                fx = N.f32(f(xview.clone()).detach())
                N.check(lib.psi_broyden_first(ws.handle, N.ptr(fx), stream), "psi_broyden_first")
                done = c_int(0)
                n = 0
                while n < threshold and not done.value:
                    fx = N.f32(f(xview.clone()).detach())
                    N.check(lib.psi_broyden_step(ws.handle, N.ptr(fx), byref(done), stream), "psi_broyden_step")
                    n += 1
                N.check(lib.psi_broyden_finish(ws.handle, N.ptr(result), byref(stats), rel, abs_, stream), "psi_broyden_finish")
            finally:
                torch.cuda.current_stream().synchronize()
                ws.close()
        trace = _trace_views(xtrace, int(stats.steps_run) + 1, numel, shape)
    return _result_dict(result, stats, rel, abs_, trace, eps, threshold)


def _as_tensor(dev_ptr: int, numel: int, device) -> torch.Tensor:
    """zero-copy fp32 view of extension-owned device memory (valid while the workspace lives)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<f4", "data": (int(dev_ptr), False), "version": 2}
    return torch.as_tensor(h, device=device)


def forward_iteration(f: Callable, z0: torch.Tensor, eps: float = 1.e-5, threshold: int = 50, keep_trace: Optional[bool] = None,
                      **kwargs) -> dict:
    """Picard iteration z ← f(z) with rel = ‖z_prev − z‖/‖z‖ (reference solver.py:301-341).

    A :class:`LayerOperator` runs as a device-resident loop (one fused layer launch + one stop-rule launch per iteration); any
    other CUDA callable is driven through the Picard step API of the extension (norms, stop rule and trace on the device, one
    Python call of ``f`` per iteration).  ``keep_trace`` fills ``xest_trace`` with every iterate as the reference does."""
    _require_cuda(z0)
    lib = N.load()
    z0c = N.f32(z0.detach())
    shape, numel = z0c.shape, z0c.numel()
    threshold = int(threshold)
    keep = KEEP_TRACE if keep_trace is None else keep_trace
    stats = N.SolveStats()
    rel = (c_double * (threshold + 2))()
    abs_ = (c_double * (threshold + 2))()
    result = torch.empty_like(z0c)
    with torch.cuda.device(z0c.device):
        stream = N.stream_ptr()
        if isinstance(f, NativeOperator) and f.op == N.OP_LAYER:
            ws = f.graph.solver(threshold)
            xtrace = torch.empty(threshold + 2, ws.stride, dtype=torch.float32, device=z0c.device) if keep else None
            f.upload()
            N.check(lib.psi_solver_picard(ws.handle, f.graph.handle, f.kind, N.ptr(z0c), N.ptr(f.aux), threshold, float(eps), N.ptr(result),
                                          byref(stats), rel, abs_, N.ptr(xtrace), stream), "psi_solver_picard")
        else:
            ws = SolverWorkspace(numel, max(threshold, 1), z0c.device)
            xtrace = torch.empty(threshold + 2, ws.stride, dtype=torch.float32, device=z0c.device) if keep else None
            try:
                N.check(lib.psi_picard_begin(ws.handle, N.ptr(z0c), threshold, float(eps), N.ptr(xtrace), stream), "psi_picard_begin")
                done = c_int(0)
                while not done.value:
                    xview = _as_tensor(lib.psi_picard_x(ws.handle), numel, z0c.device).view(shape)
                    fx = N.f32(f(xview.clone()).detach())
                    N.check(lib.psi_picard_feed(ws.handle, N.ptr(fx), byref(done), stream), "psi_picard_feed")
                N.check(lib.psi_picard_finish(ws.handle, N.ptr(result), byref(stats), rel, abs_, stream), "psi_picard_finish")
            finally:
                torch.cuda.current_stream().synchronize()
                ws.close()
    n = int(stats.f_evals)
    out = _result_dict(result, stats, list(rel)[:n], list(abs_)[:n], _trace_views(xtrace, n + 1, numel, shape), eps, threshold)
    out.pop("prot_break")
    return out


def anderson(f: Callable, x0: torch.Tensor, m: int = 2, lam: float = 1e-4, threshold: int = 50, eps: float = 1e-3,
             stop_mode: str = "rel", beta: float = 1.0, keep_trace: Optional[bool] = None, **kwargs) -> dict:
    """Anderson acceleration (reference solver.py:215-293).

    The window algebra (Gram matrix, bordered solve, mixing, norms, best-iterate bookkeeping) runs in the extension's kernels for
    every operator: a :class:`LayerOperator` as one device-resident loop, any other CUDA callable through the Anderson step API
    with one Python call of ``f`` per step."""
    if stop_mode != "rel":
        raise NotImplementedError("psi_gnn_b200.solver.anderson: only stop_mode='rel'")
    _require_cuda(x0)
    lib = N.load()
    x0c = N.f32(x0.detach())
    shape, numel = x0c.shape, x0c.numel()
    threshold = int(threshold)
    keep = KEEP_TRACE if keep_trace is None else keep_trace
    stats = N.SolveStats()
    rel = (c_double * (threshold + 1))()
    abs_ = (c_double * (threshold + 1))()
    result = torch.empty_like(x0c)
    with torch.cuda.device(x0c.device):
        stream = N.stream_ptr()
        if isinstance(f, NativeOperator) and f.op == N.OP_LAYER:
            ws = f.graph.solver(threshold)
            xtrace = torch.empty(threshold + 2, ws.stride, dtype=torch.float32, device=x0c.device) if keep else None
            f.upload()
            N.check(lib.psi_solver_anderson(ws.handle, f.graph.handle, f.kind, N.ptr(x0c), N.ptr(f.aux), int(m), float(lam), threshold,
                                            float(eps), float(beta), N.ptr(result), byref(stats), rel, abs_, N.ptr(xtrace), stream),
                    "psi_solver_anderson")
        else:
            ws = SolverWorkspace(numel, max(threshold, 1), x0c.device)
            xtrace = torch.empty(threshold + 2, ws.stride, dtype=torch.float32, device=x0c.device) if keep else None
            try:
                N.check(lib.psi_anderson_begin(ws.handle, N.ptr(x0c), int(m), float(lam), threshold, float(eps), float(beta), N.ptr(xtrace),
                                               stream), "psi_anderson_begin")
                done = c_int(0)
                while not done.value:
                    xview = _as_tensor(lib.psi_anderson_x(ws.handle), numel, x0c.device).view(shape)
                    fx = N.f32(f(xview.clone()).detach())
                    N.check(lib.psi_anderson_feed(ws.handle, N.ptr(fx), byref(done), stream), "psi_anderson_feed")
                N.check(lib.psi_anderson_finish(ws.handle, N.ptr(result), byref(stats), rel, abs_, stream), "psi_anderson_finish")
            finally:
                torch.cuda.current_stream().synchronize()
                ws.close()
    n = max(threshold - 2, 0)
    return _result_dict(result, stats, list(rel)[:n], list(abs_)[:n], _trace_views(xtrace, int(stats.steps_run) + 1, numel, shape), eps,
                        threshold)


def newton(f, z0, eps, threshold):
    raise NotImplementedError("psi_gnn_b200.solver.newton: the reference's dense-Jacobian toy solver (solver.py:343-366) is "
                              "not part of the accelerated path")
