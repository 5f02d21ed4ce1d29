"""Node-range partition of ONE large mesh over the ranks (BASELINE config 5; SURVEY §8e "one large mesh").

The reference has no counterpart (its only parallelism is graph-level ``DataParallel``); the semantics are those of the
unpartitioned solve: ``broyden`` treats the whole mesh as one vector, so every inner product and norm is a global sum.

Host logic (numpy, runs anywhere):
  * nodes are ordered by (lattice row, x) and cut into ``world`` contiguous ranges of equal size (METIS-style ranges after a
    geometric ordering: strips have ≤ 2 neighbours each and a halo of O(√N) nodes);
  * rank r keeps every matrix entry (row, col) whose row OR column it owns — exactly the edges the two aggregation
    directions of its owned nodes need — and sees the remote endpoints as ghost nodes appended after its owned nodes,
    grouped by owner;
  * ``send_index`` lists, per peer, the owned rows that are ghosts over there, in the peer's ghost order.
Device logic: ``psi_graph_set_partition`` + the partition-aware ``psi_solver_broyden`` (halo exchange with ncclSend/ncclRecv
before every operator evaluation, two small fp64 all-reduces per Broyden step), see csrc/comm.cuh.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_int32, c_int64, c_void_p
from typing import List, Optional

import numpy as np
import torch

from .synthetic import GraphData

NODE_FIELDS = ("x", "y", "sol", "prb_data", "tags", "pos", "unit_normal_vector")
EDGE_FIELDS = ("edge_attr", "a_ij")


class MeshPartition:
    """One rank's share of the mesh: ``local`` (GraphData over owned + ghost nodes) and the exchange lists."""

    def __init__(self):
        self.rank = 0
        self.world = 1
        self.n_owned = 0
        self.n_ghost = 0
        self.owned_global: np.ndarray = None        # global ids of the owned rows, in local order
        self.ghost_global: np.ndarray = None
        self.peers: List[int] = []
        self.send_counts: List[int] = []
        self.recv_counts: List[int] = []
        self.send_index: np.ndarray = None          # int32 local owned rows, concatenated per peer
        self.local: GraphData = None
        self.comm = None                            # Communicator, set by attach()


ALIGN_NODES = 2048      # lcm(4096-float reduction chunk, 10 floats per node, 128-node CTA) in nodes


def geometric_order(pos: np.ndarray) -> np.ndarray:
    """Row-major cell ordering: nodes are binned along y into rows about one mesh spacing high and sorted by (row, x).
    Consecutive indices are then horizontal neighbours, and the t-th neighbours of 32 consecutive nodes are (mostly) consecutive
    rows of the adjacent lattice row — the gathers of the layer kernels touch ≈ 13 cache lines per warp and trip instead of ≈ 29
    for a plain coordinate sort, which interleaves the jittered nodes of a lattice line at random (measured on the synthetic
    1M-node mesh; the generator's own lattice order gives 11).  Ranges of this order are horizontal strips (≤ 2 neighbours, halo
    O(√N))."""
    n = pos.shape[0]
    if n == 0:
        return np.zeros(0, np.int64)
    x, y = pos[:, 0].astype(np.float64), pos[:, 1].astype(np.float64)
    area = max((x.max() - x.min()) * (y.max() - y.min()), 1e-30)
    w = 0.75 * np.sqrt(area / n)                             # a little under the mesh spacing
    row = np.floor((y - y.min()) / w).astype(np.int64)
    return np.lexsort((x, row))


def _owner_of(data: GraphData, world: int):
    pos = data.pos.numpy()
    order = geometric_order(pos)                            # geometric ordering → contiguous ranges
    n = order.shape[0]
    if n >= 2 * ALIGN_NODES * world:
        # Cut at multiples of 2048 nodes: every rank's rows then start on a boundary of the 4096-float chunks (and 128-node blocks)
        # the reductions of the solver are organised in, so its fp32 partial sums are the very ones the unpartitioned solve of the
        # same (reordered) mesh forms — the partitioned solve retraces the single-GPU trajectory instead of a chaotic neighbour.
        bounds = np.round(np.arange(1, world) * n / world / ALIGN_NODES).astype(np.int64) * ALIGN_NODES
        rank_of_sorted = np.searchsorted(bounds, np.arange(n, dtype=np.int64), side="right")
    else:
        rank_of_sorted = np.minimum((np.arange(n, dtype=np.int64) * world) // max(n, 1), world - 1)
    owner = np.empty(n, np.int64)
    owner[order] = rank_of_sorted
    sorted_pos = np.empty(n, np.int64)
    sorted_pos[order] = np.arange(n)
    return owner, sorted_pos


def partition_mesh(data: GraphData, world: int, rank: Optional[int] = None) -> List[MeshPartition]:
    """Split a single-mesh ``GraphData`` (CPU tensors) into ``world`` parts; with ``rank`` given only that part is built."""
    owner, spos = _owner_of(data, world)
    row = data.edge_index[0].numpy()
    col = data.edge_index[1].numpy()
    orow, ocol = owner[row], owner[col]
    cut = orow != ocol
    # ghost pairs (rank that needs the node, node): the remote endpoint of every cut entry, seen from both sides
    need_rank = np.concatenate([orow[cut], ocol[cut]])
    need_node = np.concatenate([col[cut], row[cut]])
    key = np.unique(need_rank * data.num_nodes + need_node)
    g_rank, g_node = key // data.num_nodes, key % data.num_nodes
    # order ghosts by (needing rank, owner of the node, geometric position)
    o = np.lexsort((spos[g_node], owner[g_node], g_rank))
    g_rank, g_node = g_rank[o], g_node[o]
    parts = []
    for r in (range(world) if rank is None else [rank]):
        P = MeshPartition()
        P.rank, P.world = r, world
        mine = np.where(owner == r)[0]
        mine = mine[np.argsort(spos[mine], kind="stable")]
        ghosts = g_node[g_rank == r]
        P.owned_global, P.ghost_global = mine, ghosts
        P.n_owned, P.n_ghost = int(mine.shape[0]), int(ghosts.shape[0])
        local_of = -np.ones(data.num_nodes, np.int64)
        local_of[mine] = np.arange(P.n_owned)
        local_of[ghosts] = P.n_owned + np.arange(P.n_ghost)
        gown = owner[ghosts]
        # what I must send: my owned nodes that are ghosts on rank q, in q's ghost order
        send_idx = []
        for q in range(world):
            if q == r:
                continue
            recv = int((gown == q).sum())
            theirs = g_node[(g_rank == q) & (owner[g_node] == r)]
            if recv == 0 and theirs.shape[0] == 0:
                continue
            P.peers.append(q)
            P.recv_counts.append(recv)
            P.send_counts.append(int(theirs.shape[0]))
            send_idx.append(local_of[theirs])
        P.send_index = (np.concatenate(send_idx) if send_idx else np.zeros(0, np.int64)).astype(np.int32)
        # local graph: entries whose row or column is owned
        keep = (orow == r) | (ocol == r)
        loc = GraphData()
        nodes = np.concatenate([mine, ghosts])
        tn = torch.from_numpy(nodes)
        for k in NODE_FIELDS:
            v = getattr(data, k, None)
            if v is not None:
                setattr(loc, k, v[tn].contiguous())
        tk = torch.from_numpy(np.where(keep)[0])
        for k in EDGE_FIELDS:
            v = getattr(data, k, None)
            if v is not None:
                setattr(loc, k, v[tk].contiguous())
        loc.edge_index = torch.from_numpy(np.stack([local_of[row[keep]], local_of[col[keep]]])).contiguous()
        loc.num_nodes = P.n_owned + P.n_ghost
        loc.num_graphs = 1
        loc.partition = P
        P.local = loc
        parts.append(P)
    return parts


def reorder_mesh(data: GraphData) -> GraphData:
    """the mesh in the partitioner's geometric node order (a 1-rank 'partition'): the single-GPU counterpart of a partitioned solve"""
    return partition_mesh(data, 1)[0].local


class Communicator:
    """``psi_comm_t`` built from a ``torch.distributed`` process group (the 128-byte NCCL id travels through it)."""

    def __init__(self, device, group=None):
        import torch.distributed as dist
        from . import _native as N
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.handle = c_void_p()
        lib = N.load()
        buf = ctypes.create_string_buffer(128)
        with torch.cuda.device(device):
            if self.world > 1:
                if self.rank == 0:
                    N.check(lib.psi_comm_unique_id(buf), "psi_comm_unique_id")
                t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=device)
                dist.broadcast(t, src=0, group=group)
                buf = ctypes.create_string_buffer(bytes(t.cpu().tolist()), 128)
            N.check(lib.psi_comm_create(byref(self.handle), self.rank, self.world, buf), "psi_comm_create")

    def close(self):
        from . import _native as N
        if self.handle:
            N.load().psi_comm_destroy(self.handle)
            self.handle = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def attach(graph, part: MeshPartition, comm: Communicator):
    """bind the exchange lists of ``part`` to a NativeGraph built over ``part.local`` (owned + ghost rows)"""
    from . import _native as N
    n = len(part.peers)
    peers = (c_int32 * max(n, 1))(*part.peers)
    sc = (c_int64 * max(n, 1))(*part.send_counts)
    rc = (c_int64 * max(n, 1))(*part.recv_counts)
    idx = torch.from_numpy(part.send_index).to(graph.device) if part.send_index.size else None
    with torch.cuda.device(graph.device):
        N.check(N.load().psi_graph_set_partition(graph.handle, comm.handle, part.n_owned, n, peers, sc, rc, N.ptr(idx), N.stream_ptr()),
                "psi_graph_set_partition")
    graph.partition = part
    graph.comm = comm          # keep the communicator alive as long as the graph
    part.comm = comm
    if comm.world > 1 and USE_P2P:
        _open_mailboxes(graph, part, comm)
    return graph


USE_P2P = True       # peer-mapped mailboxes (CUDA IPC over NVLink); False = grouped ncclSend/ncclRecv + ncclAllReduce


def _open_mailboxes(graph, part: MeshPartition, comm: Communicator):
    """collective over the ranks of ``comm``: every rank allocates its mailbox, the IPC handles, ghost counts and ghost layouts are
    all-gathered through ``torch.distributed`` (host side, once per graph), then every rank maps its peers' blocks"""
    import torch.distributed as dist
    from . import _native as N
    lib = N.load()
    buf = ctypes.create_string_buffer(64)
    tr = c_int64(0)
    with torch.cuda.device(graph.device):
        N.check(lib.psi_part_mail_create(graph.handle, buf, byref(tr)), "psi_part_mail_create")
    recv_off, o = {}, 0
    for q, c in zip(part.peers, part.recv_counts):
        recv_off[int(q)] = o
        o += int(c)
    mine = {"handle": bytes(buf.raw), "total_recv": int(tr.value), "recv_off": recv_off}
    everyone = [None] * comm.world
    dist.all_gather_object(everyone, mine)
    handles = b"".join(e["handle"] for e in everyone)
    totals = (c_int64 * comm.world)(*[e["total_recv"] for e in everyone])
    n = len(part.peers)
    remote = (c_int64 * max(n, 1))(*[everyone[q]["recv_off"][comm.rank] for q in part.peers])
    with torch.cuda.device(graph.device):
        N.check(lib.psi_part_mail_open(graph.handle, handles, totals, remote), "psi_part_mail_open")
    dist.barrier()           # nobody stores into a block before every rank has mapped it
