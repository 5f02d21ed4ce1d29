"""Synthetic 2-D Poisson / P1-FEM mesh generator and PyG-free batch container.

The reference trains on meshes produced by gmsh + FEniCS (neither is available
offline), so every test and benchmark of this repo runs on meshes produced
here.  The generator emits exactly the per-graph arrays that the reference's
dataset reader turns into a ``torch_geometric.data.Data`` object
(reference ``dirichlet/psignn/utilities/reader.py:80-116`` and
``mixed/psignn/utilities/reader.py:84-124``):

``x, edge_index, edge_attr, a_ij, y, sol, prb_data, tags, pos`` (+
``unit_normal_vector`` for the mixed problem), with the reference's
hard-coded normalisation constants (``reader.py:73-77`` / mixed ``:74-81``).

Geometry follows ``dirichlet/dataset/build_mesh.py:57-69``: a star-shaped
domain whose boundary passes through control points at radius
``U(0.75, 1)*R``.  Physics follows ``dirichlet/dataset/extract_data.py:16-90``:
``-Δu = f`` with ``f = A(x/R-1)² + B(y/R)² + C``, Dirichlet data
``g`` = random quadratic, coefficients ``U(-10, 10)``; Dirichlet rows of the
stiffness matrix are replaced by identity rows (``bc.apply(A, b)``, not
symmetrised) and ``edge_index, a_ij = scipy.sparse.find(A)``.

Nothing here is on the GPU hot path; it is input synthesis only.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Iterable, List, Optional, Sequence

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch
from scipy.spatial import Delaunay

# reader.py:73-77 (dirichlet) -------------------------------------------------
DIRICHLET_NORM = dict(
    prb_mean=(0.0464, -0.0006), prb_std=(9.6267, 3.2935),
    dist_mean=(0.0, 0.0, 0.0655), dist_std=(0.0507, 0.0507, 0.0293),
)
# mixed/psignn/utilities/reader.py:74-81 --------------------------------------
MIXED_NORM = dict(
    prb_mean=(-0.4319, 0.0289, -0.0189), prb_std=(8.4245, 2.1942, 2.8585),
    dist_mean=(0.0, 0.0, 0.0572), dist_std=(0.0445, 0.0443, 0.0258),
    nrm_mean=(0.0007, -0.0004), nrm_std=(0.2773, 0.2959),
)

N_CONTROL = 9  # control radii of the star-shaped boundary


class GraphData:
    """Attribute bag standing in for ``torch_geometric.data.Data``/``Batch``.

    The reference only ever reads attributes off the batch object
    (``model.py:63-95,159-165,281-288``), so any object with these fields is a
    valid ``batch``.
    """

    _TENSOR_FIELDS = (
        "x", "edge_index", "edge_attr", "a_ij", "y", "sol", "prb_data", "tags",
        "pos", "unit_normal_vector", "batch", "ptr", "edge_ptr",
        # DSS-specific fields (dirichlet/dss/utilities/reader.py)
        "a_ij_norm", "b_prime", "b_prime_norm", "dss_edge_index",
    )

    def __init__(self, **kw):
        self.num_graphs = 1
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_edges(self) -> int:
        return int(self.edge_index.shape[1])

    def keys(self):
        return [k for k in self._TENSOR_FIELDS if getattr(self, k, None) is not None]

    def to(self, device, non_blocking: bool = False) -> "GraphData":
        out = GraphData()
        out.__dict__.update({k: v for k, v in self.__dict__.items() if not k.startswith("_psi")})
        for k in self.keys():
            setattr(out, k, getattr(self, k).to(device, non_blocking=non_blocking))
        return out          # non-tensor attributes (num_nodes, partition, …) travel by reference

    def pin_memory(self) -> "GraphData":
        out = GraphData()
        out.__dict__.update({k: v for k, v in self.__dict__.items() if not k.startswith("_psi")})
        for k in self.keys():
            setattr(out, k, getattr(self, k).pin_memory())
        return out

    def double(self) -> "GraphData":
        out = GraphData()
        out.__dict__.update({k: v for k, v in self.__dict__.items() if not k.startswith("_psi")})
        for k in self.keys():
            t = getattr(self, k)
            if t.is_floating_point():
                setattr(out, k, t.double())
        return out

    def nbytes(self) -> int:
        return sum(getattr(self, k).numel() * getattr(self, k).element_size() for k in self.keys())


# -----------------------------------------------------------------------------
# geometry
# -----------------------------------------------------------------------------

def _radius_fn(radii: np.ndarray):
    """Periodic piecewise-linear r(theta) through ``radii`` at equispaced angles."""
    k = len(radii)
    ang = np.linspace(0.0, 2.0 * math.pi, k + 1)
    rr = np.concatenate([radii, radii[:1]])

    def r_of(theta):
        t = np.mod(theta, 2.0 * math.pi)
        return np.interp(t, ang, rr)

    return r_of


def _boundary_ring(r_of, h: float, R: float) -> np.ndarray:
    """Points at (approximately) arclength spacing ``h`` along the boundary."""
    fine = np.linspace(0.0, 2.0 * math.pi, 20001)
    rf = r_of(fine)
    xy = np.stack([rf * np.cos(fine), rf * np.sin(fine)], 1)
    seg = np.linalg.norm(np.diff(xy, axis=0), axis=1)
    s = np.concatenate([[0.0], np.cumsum(seg)])
    m = max(8, int(round(s[-1] / h)))
    target = np.linspace(0.0, s[-1], m, endpoint=False)
    th = np.interp(target, s, fine)
    rb = r_of(th)
    return np.stack([rb * np.cos(th), rb * np.sin(th)], 1), th


def _interior_lattice(r_of, h: float, R: float, rng: np.random.Generator, jitter: float) -> np.ndarray:
    dy = h * math.sqrt(3.0) / 2.0
    ny = int(math.ceil(R / dy)) + 1
    nx = int(math.ceil(R / h)) + 1
    jj, ii = np.meshgrid(np.arange(-ny, ny + 1), np.arange(-nx, nx + 1), indexing="ij")
    px = (ii + 0.5 * (jj & 1)) * h
    py = jj * dy
    pts = np.stack([px.ravel(), py.ravel()], 1)
    pts = pts + rng.uniform(-jitter * h, jitter * h, pts.shape)
    rho = np.hypot(pts[:, 0], pts[:, 1])
    th = np.arctan2(pts[:, 1], pts[:, 0])
    keep = rho < r_of(th) - 0.72 * h
    return pts[keep]


def _triangulate(points: np.ndarray, r_of) -> np.ndarray:
    tri = Delaunay(points).simplices
    c = points[tri].mean(1)
    rho = np.hypot(c[:, 0], c[:, 1])
    th = np.arctan2(c[:, 1], c[:, 0])
    tri = tri[rho < r_of(th)]
    # drop degenerate slivers (area ~ 0)
    p = points[tri]
    area2 = np.abs((p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1])
                   - (p[:, 2, 0] - p[:, 0, 0]) * (p[:, 1, 1] - p[:, 0, 1]))
    return tri[area2 > 1e-12]


def _boundary_edges(tri: np.ndarray):
    e = np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]], 0)
    es = np.sort(e, 1)
    uniq, inv, cnt = np.unique(es, axis=0, return_inverse=True, return_counts=True)
    mask = cnt[inv.ravel()] == 1
    return e[mask]  # oriented as in the (ccw) triangles


# -----------------------------------------------------------------------------
# P1 finite elements
# -----------------------------------------------------------------------------

def _assemble_p1(points: np.ndarray, tri: np.ndarray):
    """Stiffness K = ∫∇φi·∇φj and consistent mass M for P1 triangles."""
    p = points[tri]                                    # [T,3,2]
    x, y = p[..., 0], p[..., 1]
    # make all triangles counter-clockwise
    area2 = (x[:, 1] - x[:, 0]) * (y[:, 2] - y[:, 0]) - (x[:, 2] - x[:, 0]) * (y[:, 1] - y[:, 0])
    flip = area2 < 0
    tri = tri.copy()
    tri[flip] = tri[flip][:, [0, 2, 1]]
    p = points[tri]
    x, y = p[..., 0], p[..., 1]
    area2 = np.abs(area2)
    bcoef = np.stack([y[:, 1] - y[:, 2], y[:, 2] - y[:, 0], y[:, 0] - y[:, 1]], 1)
    ccoef = np.stack([x[:, 2] - x[:, 1], x[:, 0] - x[:, 2], x[:, 1] - x[:, 0]], 1)
    ke = (bcoef[:, :, None] * bcoef[:, None, :] + ccoef[:, :, None] * ccoef[:, None, :]) / (2.0 * area2)[:, None, None]
    me = (area2 / 24.0)[:, None, None] * (np.ones((3, 3)) + np.eye(3))[None]
    rows = np.repeat(tri, 3, axis=1).ravel()
    cols = np.tile(tri, (1, 3)).ravel()
    n = points.shape[0]
    K = sp.csr_matrix((ke.ravel(), (rows, cols)), shape=(n, n))
    M = sp.csr_matrix((me.ravel(), (rows, cols)), shape=(n, n))
    return K, M, tri


def _vertex_normals(points: np.ndarray, bedges: np.ndarray) -> np.ndarray:
    """Length-weighted outward normal at boundary vertices (zero elsewhere).

    Stand-in for the FEniCS projection of the facet normal
    (``mixed/dataset/extract_data.py:119-137``).
    """
    n = np.zeros_like(points)
    t = points[bedges[:, 1]] - points[bedges[:, 0]]
    # triangles are ccw, so the outward normal of a boundary edge (a->b) is (ty, -tx)
    en = np.stack([t[:, 1], -t[:, 0]], 1)
    np.add.at(n, bedges[:, 0], 0.5 * en)
    np.add.at(n, bedges[:, 1], 0.5 * en)
    nrm = np.linalg.norm(n, axis=1, keepdims=True)
    return np.divide(n, nrm, out=np.zeros_like(n), where=nrm > 0)


# -----------------------------------------------------------------------------
# public API
# -----------------------------------------------------------------------------

def make_mesh(seed: int, h: float = 0.075, R: float = 1.0, mixed: bool = False,
              solve: bool = True, jitter: float = 0.12, dtype=torch.float32,
              keep_triangles: bool = False) -> GraphData:
    """One synthetic Poisson problem in the reference's ``Data`` layout.

    ``h=0.075, R=1`` gives ~510 nodes (configs C1–C3); ``h=0.037`` ~2 k nodes (C4).
    Large single meshes keep ``h`` and scale ``R`` so that the normalised edge
    features stay in-distribution.
    """
    rng = np.random.default_rng(seed)
    radii = rng.uniform(0.75, 1.0, N_CONTROL) * R
    r_of = _radius_fn(radii)
    ring, ring_theta = _boundary_ring(r_of, h, R)
    inner = _interior_lattice(r_of, h, R, rng, jitter)
    points = np.concatenate([inner, ring], 0)
    tri = _triangulate(points, r_of)
    # drop unreferenced points (possible near concave corners)
    used = np.zeros(points.shape[0], bool)
    used[tri.ravel()] = True
    if not used.all():
        remap = -np.ones(points.shape[0], np.int64)
        remap[used] = np.arange(int(used.sum()))
        points = points[used]
        tri = remap[tri]
    n = points.shape[0]

    K, M, tri = _assemble_p1(points, tri)
    bedges = _boundary_edges(tri)
    bnodes = np.unique(bedges.ravel())
    is_bnd = np.zeros(n, bool)
    is_bnd[bnodes] = True

    pf = rng.uniform(-10.0, 10.0, 3)
    pg = rng.uniform(-10.0, 10.0, 6)
    xs, ys = points[:, 0] / R, points[:, 1] / R
    f_val = pf[0] * (xs - 1.0) ** 2 + pf[1] * ys ** 2 + pf[2]
    g_val = pg[0] * xs * xs + pg[1] * xs * ys + pg[2] * ys * ys + pg[3] * xs + pg[4] * ys + pg[5]

    if mixed:
        # four boundary arcs, alternately Dirichlet / Neumann (mixed/dataset/build_mesh.py:78-106)
        th = np.mod(np.arctan2(points[:, 1], points[:, 0]), 2.0 * math.pi)
        quarter = np.minimum((th / (0.5 * math.pi)).astype(np.int64), 3)
        sense = int(rng.integers(0, 2))
        dir_q = (quarter % 2) == (0 if sense == 1 else 1)
        is_dir = is_bnd & dir_q
        # arc end points belong to the Dirichlet curve in gmsh's physical groups: make sure
        # every Neumann run is closed by Dirichlet nodes (purely cosmetic here)
        is_neu = is_bnd & ~is_dir
    else:
        is_dir = is_bnd
        is_neu = np.zeros(n, bool)

    b = M @ f_val
    dir_idx = np.where(is_dir)[0]
    # bc.apply(A, b): identity rows, b_i = g_i (columns untouched)
    A = K.tocsr(copy=True)
    keep_row = np.ones(n)
    keep_row[dir_idx] = 0.0
    A = sp.diags(keep_row) @ A + sp.diags(1.0 - keep_row)
    A = A.tocsr()
    A.eliminate_zeros()
    b = b.copy()
    b[dir_idx] = g_val[dir_idx]

    sol = None
    if solve:
        sol = spla.spsolve(A.tocsc(), b)

    row, col, val = sp.find(A)
    order = np.lexsort((col, row))
    row, col, val = row[order], col[order], val[order]
    d = points[row] - points[col]
    dist = np.concatenate([d, np.linalg.norm(d, axis=1, keepdims=True)], 1)

    if mixed:
        nrm = MIXED_NORM
        tags = np.zeros((n, 3))
        tags[:, 0] = 1.0
        tags[is_bnd, 0] = 0.0
        tags[is_neu, 2] = 1.0
        tags[is_dir, 1] = 1.0
        prb = np.zeros((n, 3))
        prb[:, 0] = f_val
        prb[is_neu, 2] = f_val[is_neu]
        prb[is_bnd, 0] = 0.0
        prb[is_dir, 1] = g_val[is_dir]
        unv = _vertex_normals(points, bedges)
        unv_n = (unv - np.asarray(nrm["nrm_mean"])) / np.asarray(nrm["nrm_std"])
    else:
        nrm = DIRICHLET_NORM
        tags = is_dir.astype(np.float64).reshape(-1, 1)
        prb = np.stack([np.where(is_dir, 0.0, f_val), np.where(is_dir, g_val, 0.0)], 1)
        unv_n = None

    prb_n = (prb - np.asarray(nrm["prb_mean"])) / np.asarray(nrm["prb_std"])
    dist_n = (dist - np.asarray(nrm["dist_mean"])) / np.asarray(nrm["dist_std"])

    def t(a):
        return torch.tensor(np.ascontiguousarray(a), dtype=dtype)

    y = t(b.reshape(-1, 1))
    x = torch.zeros_like(y)
    x[dir_idx] = y[dir_idx]
    data = GraphData(
        x=x,
        edge_index=torch.tensor(np.stack([row, col]), dtype=torch.long),
        edge_attr=t(dist_n), a_ij=t(val.reshape(-1, 1)), y=y,
        sol=t(sol.reshape(-1, 1)) if sol is not None else torch.zeros_like(y),
        prb_data=t(prb_n), tags=t(tags), pos=t(points),
    )
    if mixed:
        data.unit_normal_vector = t(unv_n)
    data.num_nodes = n
    if keep_triangles:
        data.triangles = tri
    return data


# dirichlet/dss/utilities/reader.py:63-67 ---------------------------------------
DSS_NORM = dict(aij_mean=-0.5838, aij_std=0.0924, b_mean=(0.0002, 0.1435, -0.0006), b_std=(0.0507, 0.3506, 3.2935))


def to_dss(data: GraphData) -> GraphData:
    """The same problem in the DSS reader's layout (dirichlet/dataset/generate_data.py:100-128 ``add_dss_variable`` +
    dirichlet/dss/utilities/reader.py:69-93): A' = A with the diagonal removed (Dirichlet rows become empty),
    b' = [b, 0, 0] with Dirichlet rows [0, 1, g]; ``x = sol``; edge feature = normalised a_ij."""
    row, col = data.edge_index
    keep = row != col
    ei = data.edge_index[:, keep].contiguous()
    a = data.a_ij[keep].contiguous()
    is_dir = data.tags.reshape(-1) == 1
    b = data.y.reshape(-1)
    bp = torch.stack([torch.where(is_dir, torch.zeros_like(b), b), is_dir.to(b.dtype), torch.where(is_dir, b, torch.zeros_like(b))], 1)
    out = GraphData(x=data.sol.clone(), edge_index=ei, a_ij=a, a_ij_norm=(a - DSS_NORM["aij_mean"]) / DSS_NORM["aij_std"],
                    b_prime=bp, b_prime_norm=(bp - torch.tensor(DSS_NORM["b_mean"], dtype=b.dtype)) / torch.tensor(DSS_NORM["b_std"], dtype=b.dtype),
                    pos=data.pos, tags=data.tags, sol=data.sol)
    out.num_nodes = data.num_nodes
    out.num_graphs = getattr(data, "num_graphs", 1)
    return out


def collate(graphs: Sequence[GraphData]) -> GraphData:
    """``Batch.from_data_list`` semantics: concatenate on dim 0, offset ``edge_index``."""
    graphs = list(graphs)
    out = GraphData()
    offs = np.cumsum([0] + [g.num_nodes for g in graphs])
    eoffs = np.cumsum([0] + [g.num_edges for g in graphs])
    for k in graphs[0].keys():
        if k == "edge_index":
            out.edge_index = torch.cat([g.edge_index + int(o) for g, o in zip(graphs, offs[:-1])], 1)
        elif k in ("batch", "ptr", "edge_ptr"):
            continue
        else:
            setattr(out, k, torch.cat([getattr(g, k) for g in graphs], 0))
    out.batch = torch.cat([torch.full((g.num_nodes,), i, dtype=torch.long) for i, g in enumerate(graphs)])
    out.ptr = torch.tensor(offs, dtype=torch.long)
    out.edge_ptr = torch.tensor(eoffs, dtype=torch.long)
    out.num_nodes = int(offs[-1])
    out.num_graphs = len(graphs)
    return out


def make_batch(num_graphs: int, seed0: int = 0, **kw) -> GraphData:
    return collate([make_mesh(seed0 + i, **kw) for i in range(num_graphs)])


def make_large_mesh(num_nodes: int, seed: int = 0, h: float = 0.075, mixed: bool = False,
                    solve: bool = False, dtype=torch.float32) -> GraphData:
    """Single mesh with ≈ ``num_nodes`` nodes: keep ``h``, scale the domain radius."""
    # nodes ≈ area / (h² √3/2); mean r² of U(0.75,1)² ≈ 0.77
    R = math.sqrt(num_nodes * h * h * math.sqrt(3.0) / 2.0 / (math.pi * 0.77))
    return make_mesh(seed, h=h, R=R, mixed=mixed, solve=solve, dtype=dtype)


def split_graphs(batch: GraphData, parts: int) -> List[GraphData]:
    """Split a collated batch into ``parts`` contiguous groups of graphs balanced by node count.

    Mirrors what PyG ``DataParallel.scatter`` does with the data list
    (reference call site ``dirichlet/psignn/main.py:106``).
    """
    ptr = batch.ptr.numpy()
    eptr = batch.edge_ptr.numpy()
    g = batch.num_graphs
    total = ptr[-1]
    bounds = [0]
    for p in range(1, parts):
        tgt = total * p / parts
        j = int(np.argmin(np.abs(ptr - tgt)))
        j = max(bounds[-1] + 1, min(j, g - (parts - p)))
        bounds.append(j)
    bounds.append(g)
    outs = []
    for p in range(parts):
        g0, g1 = bounds[p], bounds[p + 1]
        n0, n1 = int(ptr[g0]), int(ptr[g1])
        e0, e1 = int(eptr[g0]), int(eptr[g1])
        sub = GraphData()
        for k in batch.keys():
            v = getattr(batch, k)
            if k == "edge_index":
                sub.edge_index = v[:, e0:e1] - n0
            elif k in ("edge_attr", "a_ij", "a_ij_norm"):
                setattr(sub, k, v[e0:e1])
            elif k == "ptr":
                sub.ptr = v[g0:g1 + 1] - n0
            elif k == "edge_ptr":
                sub.edge_ptr = v[g0:g1 + 1] - e0
            elif k == "batch":
                sub.batch = v[n0:n1] - g0
            elif k == "dss_edge_index":
                continue
            else:
                setattr(sub, k, v[n0:n1])
        sub.num_nodes = n1 - n0
        sub.num_graphs = g1 - g0
        outs.append(sub)
    return outs
