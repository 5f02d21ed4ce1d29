"""Packing of ``nn.Module`` parameters into the layer-constant block of ``csrc/weights.cuh``.

The block is a flat fp32 vector in the exact field order of ``struct LayerWeights``; fields a layer
variant does not use stay zero.  Key names are the reference's ``state_dict`` keys
(dirichlet/psignn/model.py:263-277, mixed/psignn/model.py:196-214, dirichlet/dss/model.py:59-104,
dirichlet/dsgps/model.py:48-131), so shipped checkpoints pack without renaming.
"""
from __future__ import annotations

from typing import Dict, Mapping, Optional

import torch

D = 10
# (field name, number of floats) in struct order
_EDGE = [("W1i", D * D), ("W1j", D * D), ("W1a", D * 3), ("b1", D), ("W2", D * D), ("b2", D)]
FIELDS = (
    [("to." + n, s) for n, s in _EDGE] + [("from." + n, s) for n, s in _EDGE] + [("neu." + n, s) for n, s in _EDGE]
    + [("gate_w", 33), ("gate_b", 1),
       ("up_W1", D * 33), ("up_b1", D), ("up_W2", D * D), ("up_b2", D),
       ("un_W1", D * 25), ("un_b1", D), ("un_W2", D * D), ("un_b2", D),
       ("ln_g", D), ("ln_b", D),
       ("gz_W", D * 33), ("gz_b", D), ("gr_W", D * 33), ("gr_b", D), ("gc_W", D * 33), ("gc_b", D),
       ("enc_W1", D), ("enc_b1", D), ("enc_W2", D * D), ("enc_b2", D),
       ("dec_W1", D * D), ("dec_b1", D), ("dec_W2", D), ("dec_b2", 1),
       ("dss_alpha", 1), ("pad", 4)]
)
OFFSETS: Dict[str, int] = {}
_o = 0
for _n, _s in FIELDS:
    OFFSETS[_n] = _o
    _o += _s
MAIN_FLOATS = _o           # 3200 = sizeof(LayerWeights) / 4

# struct LayerWeightsT: transposed ([input][output]) copies for the packed-FMA forward kernels — (source field, rows=out, pitch, inputs)
_EDGE_T = [("W1i", D, D, D), ("W1j", D, D, D), ("W1a", D, 3, 3), ("W2", D, D, D)]
TRANSPOSED = (
    [("%s.%s" % (m, n), r, p, c) for m in ("to", "from", "neu") for n, r, p, c in _EDGE_T]
    + [("up_W1", D, 33, 33), ("up_W2", D, D, D), ("un_W1", D, 25, 25), ("un_W2", D, D, D),
       ("gz_W", D, 33, 33), ("gr_W", D, 33, 33), ("gc_W", D, 33, 33)]
)
TOTAL_FLOATS = MAIN_FLOATS + sum(c * r for _, r, _, c in TRANSPOSED)      # 5960; checked against psi_weights_floats() at load time


def finish(blob: torch.Tensor) -> torch.Tensor:
    """fill the LayerWeightsT tail of a packed block from its LayerWeights head (generic: works for every layer kind)"""
    o = MAIN_FLOATS
    for name, rows, pitch, cols in TRANSPOSED:
        src = blob[OFFSETS[name]:OFFSETS[name] + rows * pitch].view(rows, pitch)[:, :cols]
        blob[o:o + cols * rows] = src.t().reshape(-1)
        o += cols * rows
    return blob


def _put(blob: torch.Tensor, name: str, t: torch.Tensor, rows: Optional[int] = None, width: Optional[int] = None):
    """write ``t`` ([rows, k] with k <= width, or a vector) into field ``name`` (row pitch ``width``)."""
    off = OFFSETS[name]
    t = t.detach().to(blob.dtype)
    if rows is None:
        blob[off:off + t.numel()] = t.reshape(-1)
    else:
        view = blob[off:off + rows * width].view(rows, width)
        view[:, :t.shape[1]] = t


def _edge(blob: torch.Tensor, slot: str, P: Mapping[str, torch.Tensor], prefix: str):
    """``Phi`` edge MLP: Linear(2d+A, d) → ReLU → Linear(d, d); first layer split by input block."""
    W1 = P[prefix + ".0.weight"]
    _put(blob, slot + ".W1i", W1[:, :D], D, D)
    _put(blob, slot + ".W1j", W1[:, D:2 * D], D, D)
    _put(blob, slot + ".W1a", W1[:, 2 * D:], D, 3)
    _put(blob, slot + ".b1", P[prefix + ".0.bias"])
    _put(blob, slot + ".W2", P[prefix + ".2.weight"], D, D)
    _put(blob, slot + ".b2", P[prefix + ".2.bias"])


def _autoencoder(blob: torch.Tensor, P: Mapping[str, torch.Tensor], prefix: str = "autoencoder"):
    e, d = prefix + ".encoder.mlp.mlp", prefix + ".decoder.mlp.mlp"
    if e + ".0.weight" in P:
        _put(blob, "enc_W1", P[e + ".0.weight"].reshape(-1))
        _put(blob, "enc_b1", P[e + ".0.bias"])
        _put(blob, "enc_W2", P[e + ".2.weight"], D, D)
        _put(blob, "enc_b2", P[e + ".2.bias"])
    if d + ".0.weight" in P:
        _decoder(blob, P, d)


def _decoder(blob: torch.Tensor, P: Mapping[str, torch.Tensor], d: str):
    _put(blob, "dec_W1", P[d + ".0.weight"], D, D)
    _put(blob, "dec_b1", P[d + ".0.bias"])
    _put(blob, "dec_W2", P[d + ".2.weight"].reshape(-1))
    _put(blob, "dec_b2", P[d + ".2.bias"])


def _check(P: Mapping[str, torch.Tensor], key: str):
    if P[key].shape[0] != D:
        raise RuntimeError("psi_gnn_b200: the native kernels are built for latent_dim = 10 (got %d)" % P[key].shape[0])


def pack_psignn(P: Mapping[str, torch.Tensor], mixed: bool, device, f_prefix: str = "deqdss.f", layer: int = 0) -> torch.Tensor:
    """PSI-GNN layer (+ autoencoder if present in ``P``)."""
    blob = torch.zeros(TOTAL_FLOATS, dtype=torch.float32, device=device)
    f = f_prefix
    _check(P, f"{f}.phi_to_list.{layer}.mlp.mlp.2.weight")
    _edge(blob, "to", P, f"{f}.phi_to_list.{layer}.mlp.mlp")
    _edge(blob, "from", P, f"{f}.phi_from_list.{layer}.mlp.mlp")
    _put(blob, "gate_w", P[f"{f}.alpha.0.weight"].reshape(-1))
    _put(blob, "gate_b", P[f"{f}.alpha.0.bias"])
    _put(blob, "up_W1", P[f"{f}.update_list.{layer}.mlp.0.weight"], D, 33)
    _put(blob, "up_b1", P[f"{f}.update_list.{layer}.mlp.0.bias"])
    _put(blob, "up_W2", P[f"{f}.update_list.{layer}.mlp.2.weight"], D, D)
    _put(blob, "up_b2", P[f"{f}.update_list.{layer}.mlp.2.bias"])
    _put(blob, "ln_g", P[f"{f}.laynorm.weight"])
    _put(blob, "ln_b", P[f"{f}.laynorm.bias"])
    if mixed:
        _edge(blob, "neu", P, f"{f}.phi_neumann.mlp.mlp")
        _put(blob, "un_W1", P[f"{f}.update_neumann.mlp.0.weight"], D, 25)
        _put(blob, "un_b1", P[f"{f}.update_neumann.mlp.0.bias"])
        _put(blob, "un_W2", P[f"{f}.update_neumann.mlp.2.weight"], D, D)
        _put(blob, "un_b2", P[f"{f}.update_neumann.mlp.2.bias"])
    _autoencoder(blob, P)
    return finish(blob)


def pack_dss(P: Mapping[str, torch.Tensor], k: int, alpha: float, device) -> torch.Tensor:
    """k-th DSS layer: Phi_to/Phi_from (edge attr = normalised a_ij), Psi, Decoder_k (dirichlet/dss/model.py:83-104)."""
    blob = torch.zeros(TOTAL_FLOATS, dtype=torch.float32, device=device)
    _check(P, f"phi_to_list.{k}.mlp.mlp.2.weight")
    _edge(blob, "to", P, f"phi_to_list.{k}.mlp.mlp")
    _edge(blob, "from", P, f"phi_from_list.{k}.mlp.mlp")
    _put(blob, "up_W1", P[f"psi_list.{k}.mlp.mlp.0.weight"], D, 33)
    _put(blob, "up_b1", P[f"psi_list.{k}.mlp.mlp.0.bias"])
    _put(blob, "up_W2", P[f"psi_list.{k}.mlp.mlp.2.weight"], D, D)
    _put(blob, "up_b2", P[f"psi_list.{k}.mlp.mlp.2.bias"])
    if f"decoder_list.{k}.mlp.mlp.0.weight" in P:
        _decoder(blob, P, f"decoder_list.{k}.mlp.mlp")
    blob[OFFSETS["dss_alpha"]] = float(alpha)
    return finish(blob)


def pack_dsgps(P: Mapping[str, torch.Tensor], device) -> torch.Tensor:
    """DSGPS recurrent step: Phi_to/Phi_from + z_k, r_k, correction gates (dirichlet/dsgps/model.py:110-131); the mixed family adds
    phi_neumann / update_neumann and a 3-column second member (mixed/dsgps/model.py:37-47)."""
    blob = torch.zeros(TOTAL_FLOATS, dtype=torch.float32, device=device)
    _check(P, "phi_to.mlp.mlp.2.weight")
    _edge(blob, "to", P, "phi_to.mlp.mlp")
    _edge(blob, "from", P, "phi_from.mlp.mlp")
    for slot, key in (("gz", "z_k"), ("gr", "r_k"), ("gc", "correction")):
        _put(blob, slot + "_W", P[f"{key}.mlp.0.weight"], D, 33)
        _put(blob, slot + "_b", P[f"{key}.mlp.0.bias"])
    if "phi_neumann.mlp.mlp.0.weight" in P:
        _edge(blob, "neu", P, "phi_neumann.mlp.mlp")
        _put(blob, "un_W1", P["update_neumann.mlp.0.weight"], D, 25)
        _put(blob, "un_b1", P["update_neumann.mlp.0.bias"])
        _put(blob, "un_W2", P["update_neumann.mlp.2.weight"], D, D)
        _put(blob, "un_b2", P["update_neumann.mlp.2.bias"])
    _autoencoder(blob, P)
    return finish(blob)


_EPOCH = [0]        # bumped by invalidate(); part of every module's pack-cache key


# ---- parameter gradients (csrc/pgrad.cuh) -----------------------------------------------------------------------------------
# record layout of one node, mirrored from the enum of pgrad.cuh (checked against psi_pgrad_layout() when the table is built)
PG = dict(ONE=0, DEG=1, C=4, CN=37, YB=62, RHAT=72, MB=82, HID=92, TB=102, SB=112, EDGE=113, ACC=323, MBN=353, HIDN=363, TBN=373, REC=383)
_PG_ORDER = ("ONE", "DEG", "C", "CN", "YB", "RHAT", "MB", "HID", "TB", "SB", "EDGE", "ACC", "MBN", "HIDN", "TBN", "REC")


def grad_table(mixed: bool):
    """(dst, ty, tx) int32 lists: gradient of the float at offset dst of the packed block = Σ_nodes record[ty]·record[tx].
    Derivation: csrc/pgrad.cuh (per-node cotangents of vjp.cuh phase A times forward intermediates; the W1j blocks regrouped by
    source node).  Only the fields a layer kind owns are listed."""
    dst, ty, tx = [], [], []

    def add(d, y, x):
        dst.append(d); ty.append(y); tx.append(x)

    for w, name in enumerate(("to", "from", "neu")[: 3 if mixed else 2]):
        eb = PG["EDGE"] + 70 * w
        for o in range(D):
            for i in range(D):
                add(OFFSETS[name + ".W1i"] + o * D + i, eb + 20 + o, PG["C"] + i)              # zs_X ⊗ h_dst
                add(OFFSETS[name + ".W1j"] + o * D + i, PG["ACC"] + 10 * w + o, PG["C"] + i)    # acc_X ⊗ h_src (node as source)
                add(OFFSETS[name + ".W2"] + o * D + i, eb + o, eb + 10 + i)                     # m̄X ⊗ S_X
            for c in range(3):
                add(OFFSETS[name + ".W1a"] + o * 3 + c, eb + 30 + o, eb + 40 + 3 * o + c)       # S̄X[o]·Σ_e mask_e[o] a_e[c]
            add(OFFSETS[name + ".b1"] + o, eb + 20 + o, PG["ONE"])
            add(OFFSETS[name + ".b2"] + o, eb + o, PG["DEG"] + w)
    width = 33 if mixed else 32
    for i in range(width):
        add(OFFSETS["gate_w"] + i, PG["SB"], PG["C"] + i)
    add(OFFSETS["gate_b"], PG["SB"], PG["ONE"])
    for o in range(D):
        for i in range(width):
            add(OFFSETS["up_W1"] + o * 33 + i, PG["TB"] + o, PG["C"] + i)
        add(OFFSETS["up_b1"] + o, PG["TB"] + o, PG["ONE"])
        for i in range(D):
            add(OFFSETS["up_W2"] + o * D + i, PG["MB"] + o, PG["HID"] + i)
        add(OFFSETS["up_b2"] + o, PG["MB"] + o, PG["ONE"])
        add(OFFSETS["ln_g"] + o, PG["YB"] + o, PG["RHAT"] + o)
        add(OFFSETS["ln_b"] + o, PG["YB"] + o, PG["ONE"])
        if mixed:
            for i in range(25):
                add(OFFSETS["un_W1"] + o * 25 + i, PG["TBN"] + o, PG["CN"] + i)
            add(OFFSETS["un_b1"] + o, PG["TBN"] + o, PG["ONE"])
            for i in range(D):
                add(OFFSETS["un_W2"] + o * D + i, PG["MBN"] + o, PG["HIDN"] + i)
            add(OFFSETS["un_b2"] + o, PG["MBN"] + o, PG["ONE"])
    return dst, ty, tx


_TABLES = {}


def grad_table_device(mixed: bool, device):
    """the table as three int32 device tensors (cached per device); verifies the record layout against the extension once"""
    key = (bool(mixed), str(device))
    if key not in _TABLES:
        import ctypes
        from . import _native as N
        lay = (ctypes.c_int32 * 16)()
        N.check(N.load().psi_pgrad_layout(lay), "psi_pgrad_layout")
        if [int(v) for v in lay] != [PG[k] for k in _PG_ORDER]:
            raise RuntimeError("psi_gnn_b200: weights.py parameter-gradient record layout does not match the extension")
        _TABLES[key] = tuple(torch.tensor(t, dtype=torch.int32, device=device) for t in grad_table(mixed))
    return _TABLES[key]


def unpack_psignn_grads(flat: torch.Tensor, names, mixed: bool, f_prefix: str = "deqdss.f", layer: int = 0) -> Dict[str, torch.Tensor]:
    """inverse of :func:`pack_psignn` for a gradient in block layout: ``{state_dict key: gradient tensor}`` for the keys in ``names``"""
    f = f_prefix
    out: Dict[str, torch.Tensor] = {}

    def field(name, rows=None, width=None, cols=None):
        off = OFFSETS[name]
        if rows is None:
            return flat[off:off + (cols or 1)]
        return flat[off:off + rows * width].view(rows, width)[:, :cols if cols is not None else width]

    def edge(slot, prefix):
        out[prefix + ".0.weight"] = torch.cat([field(slot + ".W1i", D, D), field(slot + ".W1j", D, D), field(slot + ".W1a", D, 3)], dim=1)
        out[prefix + ".0.bias"] = field(slot + ".b1", cols=D)
        out[prefix + ".2.weight"] = field(slot + ".W2", D, D)
        out[prefix + ".2.bias"] = field(slot + ".b2", cols=D)

    width = 33 if mixed else 32
    edge("to", f"{f}.phi_to_list.{layer}.mlp.mlp")
    edge("from", f"{f}.phi_from_list.{layer}.mlp.mlp")
    out[f"{f}.alpha.0.weight"] = field("gate_w", cols=width).reshape(1, width)
    out[f"{f}.alpha.0.bias"] = field("gate_b", cols=1)
    out[f"{f}.update_list.{layer}.mlp.0.weight"] = field("up_W1", D, 33, width)
    out[f"{f}.update_list.{layer}.mlp.0.bias"] = field("up_b1", cols=D)
    out[f"{f}.update_list.{layer}.mlp.2.weight"] = field("up_W2", D, D)
    out[f"{f}.update_list.{layer}.mlp.2.bias"] = field("up_b2", cols=D)
    out[f"{f}.laynorm.weight"] = field("ln_g", cols=D)
    out[f"{f}.laynorm.bias"] = field("ln_b", cols=D)
    if mixed:
        edge("neu", f"{f}.phi_neumann.mlp.mlp")
        out[f"{f}.update_neumann.mlp.0.weight"] = field("un_W1", D, 25)
        out[f"{f}.update_neumann.mlp.0.bias"] = field("un_b1", cols=D)
        out[f"{f}.update_neumann.mlp.2.weight"] = field("un_W2", D, D)
        out[f"{f}.update_neumann.mlp.2.bias"] = field("un_b2", cols=D)
    return {k: out[k] for k in names if k in out}


# ---- baselines (csrc/baseline_bwd.cuh): same record layout, other meanings of the node slots ---------------------------------------
def baseline_grad_table(kind: int):
    """(dst, ty, tx) of one unrolled DSS (kind 2) / DSGPS (3) / mixed DSGPS (4) layer; record slots as documented in baseline_bwd.cuh:
    TB = t̄ (DSS) or ā (z gate), MB = α·ȳ (DSS) or b̄ (r gate), YB = ē (correction), HID = Ψ hidden, RHAT = r ⊙ h."""
    dss, mixed = kind == 2, kind == 4
    attr = 1 if dss else 3
    dst, ty, tx = [], [], []

    def add(d, y, x):
        dst.append(d); ty.append(y); tx.append(x)

    for w, name in enumerate(("to", "from", "neu")[: 3 if mixed else 2]):
        eb = PG["EDGE"] + 70 * w
        for o in range(D):
            for i in range(D):
                add(OFFSETS[name + ".W1i"] + o * D + i, eb + 20 + o, PG["C"] + i)
                add(OFFSETS[name + ".W1j"] + o * D + i, PG["ACC"] + 10 * w + o, PG["C"] + i)
                add(OFFSETS[name + ".W2"] + o * D + i, eb + o, eb + 10 + i)
            for c in range(attr):
                add(OFFSETS[name + ".W1a"] + o * 3 + c, eb + 30 + o, eb + 40 + 3 * o + c)
            add(OFFSETS[name + ".b1"] + o, eb + 20 + o, PG["ONE"])
            add(OFFSETS[name + ".b2"] + o, eb + o, PG["DEG"] + w)
    if dss:
        for o in range(D):
            for i in range(33):
                add(OFFSETS["up_W1"] + o * 33 + i, PG["TB"] + o, PG["C"] + i)
            add(OFFSETS["up_b1"] + o, PG["TB"] + o, PG["ONE"])
            for i in range(D):
                add(OFFSETS["up_W2"] + o * D + i, PG["MB"] + o, PG["HID"] + i)
            add(OFFSETS["up_b2"] + o, PG["MB"] + o, PG["ONE"])
        return dst, ty, tx
    width = 33 if mixed else 32
    for o in range(D):
        for i in range(width):
            add(OFFSETS["gz_W"] + o * 33 + i, PG["TB"] + o, PG["C"] + i)
            add(OFFSETS["gr_W"] + o * 33 + i, PG["MB"] + o, PG["C"] + i)
            add(OFFSETS["gc_W"] + o * 33 + i, PG["YB"] + o, (PG["RHAT"] + i) if i < D else (PG["C"] + i))
        add(OFFSETS["gz_b"] + o, PG["TB"] + o, PG["ONE"])
        add(OFFSETS["gr_b"] + o, PG["MB"] + o, PG["ONE"])
        add(OFFSETS["gc_b"] + o, PG["YB"] + o, PG["ONE"])
        if mixed:
            for i in range(25):
                add(OFFSETS["un_W1"] + o * 25 + i, PG["TBN"] + o, PG["CN"] + i)
            add(OFFSETS["un_b1"] + o, PG["TBN"] + o, PG["ONE"])
            for i in range(D):
                add(OFFSETS["un_W2"] + o * D + i, PG["MBN"] + o, PG["HIDN"] + i)
            add(OFFSETS["un_b2"] + o, PG["MBN"] + o, PG["ONE"])
    return dst, ty, tx


def baseline_grad_table_device(kind: int, device):
    key = ("baseline", int(kind), str(device))
    if key not in _TABLES:
        grad_table_device(False, device)            # verifies the record layout against the extension
        _TABLES[key] = tuple(torch.tensor(t, dtype=torch.int32, device=device) for t in baseline_grad_table(kind))
    return _TABLES[key]


def _field(flat, name, rows=None, width=None, cols=None):
    off = OFFSETS[name]
    if rows is None:
        return flat[off:off + (cols or 1)]
    return flat[off:off + rows * width].view(rows, width)[:, :cols if cols is not None else width]


def _edge_grads(flat, slot, prefix, attr, out):
    out[prefix + ".0.weight"] = torch.cat([_field(flat, slot + ".W1i", D, D), _field(flat, slot + ".W1j", D, D), _field(flat, slot + ".W1a", D, 3, attr)], dim=1)
    out[prefix + ".0.bias"] = _field(flat, slot + ".b1", cols=D)
    out[prefix + ".2.weight"] = _field(flat, slot + ".W2", D, D)
    out[prefix + ".2.bias"] = _field(flat, slot + ".b2", cols=D)


def unpack_dss_grads(flat: torch.Tensor, k: int) -> Dict[str, torch.Tensor]:
    """inverse of :func:`pack_dss` for a gradient in block layout (the decoder of the block is not a layer parameter)"""
    out: Dict[str, torch.Tensor] = {}
    _edge_grads(flat, "to", f"phi_to_list.{k}.mlp.mlp", 1, out)
    _edge_grads(flat, "from", f"phi_from_list.{k}.mlp.mlp", 1, out)
    out[f"psi_list.{k}.mlp.mlp.0.weight"] = _field(flat, "up_W1", D, 33)
    out[f"psi_list.{k}.mlp.mlp.0.bias"] = _field(flat, "up_b1", cols=D)
    out[f"psi_list.{k}.mlp.mlp.2.weight"] = _field(flat, "up_W2", D, D)
    out[f"psi_list.{k}.mlp.mlp.2.bias"] = _field(flat, "up_b2", cols=D)
    return out


def unpack_dsgps_grads(flat: torch.Tensor, mixed: bool) -> Dict[str, torch.Tensor]:
    """inverse of :func:`pack_dsgps` for the recurrent step's parameters"""
    out: Dict[str, torch.Tensor] = {}
    width = 33 if mixed else 32
    _edge_grads(flat, "to", "phi_to.mlp.mlp", 3, out)
    _edge_grads(flat, "from", "phi_from.mlp.mlp", 3, out)
    for slot, key in (("gz", "z_k"), ("gr", "r_k"), ("gc", "correction")):
        out[f"{key}.mlp.0.weight"] = _field(flat, slot + "_W", D, 33, width)
        out[f"{key}.mlp.0.bias"] = _field(flat, slot + "_b", cols=D)
    if mixed:
        _edge_grads(flat, "neu", "phi_neumann.mlp.mlp", 3, out)
        out["update_neumann.mlp.0.weight"] = _field(flat, "un_W1", D, 25)
        out["update_neumann.mlp.0.bias"] = _field(flat, "un_b1", cols=D)
        out["update_neumann.mlp.2.weight"] = _field(flat, "un_W2", D, D)
        out["update_neumann.mlp.2.bias"] = _field(flat, "un_b2", cols=D)
    return out


def named_tensors(module: torch.nn.Module, prefix: str = "") -> Dict[str, torch.Tensor]:
    """parameters of ``module`` keyed like its ``state_dict`` (optionally under ``prefix``)."""
    return {(prefix + k): v for k, v in module.named_parameters()}


def version_key(P: Mapping[str, torch.Tensor]):
    """changes whenever a parameter is replaced or modified in place (optimizer step, load_state_dict).

    Writes through ``p.data`` (``p.data.copy_()``, ``p.data.mul_()`` …) do not bump ``Tensor._version`` and are therefore NOT
    seen: call :func:`invalidate` after such an update (documented in INTEGRATION.md)."""
    return (_EPOCH[0],) + tuple((k, v.data_ptr(), v._version) for k, v in P.items())


# ---- the process-wide constant bank ---------------------------------------------------------------------
_UPLOADED = {}      # device index -> key of the block currently in __constant__ memory
_LAST_STREAM = {}   # device index -> CUDA stream the bank was last written / used on
_SERIAL = [0]


def next_serial() -> int:
    """unique identity of a packed block (object ids / data pointers can be recycled by the allocator, a counter cannot)"""
    _SERIAL[0] += 1
    return _SERIAL[0]


def upload(blob: torch.Tensor, key) -> None:
    """stream-ordered copy of a packed block into the layer-constant bank unless ``key`` is already resident."""
    from . import _native as N
    dev = blob.device.index if blob.device.index is not None else torch.cuda.current_device()
    if _UPLOADED.get(dev) == key:
        return
    lib = N.load()
    if lib.psi_weights_floats() != TOTAL_FLOATS:
        raise RuntimeError("psi_gnn_b200: weights.py layout (%d floats) does not match the extension (%d)"
                           % (TOTAL_FLOATS, lib.psi_weights_floats()))
    with torch.cuda.device(blob.device):
        # ONE constant bank per device: uploads and the kernels reading it are ordered on a single stream.  A model that
        # moves to another stream (or a second model on another stream) must not overwrite the bank under kernels still in
        # flight on the previous stream: drain the device first (rare path; one stream per device is the supported use).
        cur = N.stream_ptr()
        if _LAST_STREAM.get(dev, cur) != cur:
            torch.cuda.synchronize(blob.device)
        _LAST_STREAM[dev] = cur
        N.check(lib.psi_weights_upload(N.ptr(blob), TOTAL_FLOATS, cur), "psi_weights_upload")
    _UPLOADED[dev] = key


def mark_resident(device, key) -> None:
    """record that a native call left the block identified by ``key`` in the constant bank"""
    device = torch.device(device)
    _UPLOADED[device.index if device.index is not None else torch.cuda.current_device()] = key


def invalidate() -> None:
    """forget what the constant bank holds (and make every module re-pack): required after ``.data`` writes to parameters"""
    _UPLOADED.clear()
    _SERIAL[0] += 1
    _EPOCH[0] += 1

