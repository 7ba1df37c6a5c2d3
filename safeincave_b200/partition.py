"""Cell partition of a tetrahedral mesh across GPUs (one process per GPU) and the halo plan.

The reference partitions by cells through DOLFINx/PETSc over MPI (Grid.py:275-279, ghost updates at
MomentumEquation.py:915-922).  Here: cells are already ordered along a Morton curve, rank r takes the
r-th contiguous chunk (no METIS needed), nodes referenced by more than one chunk are DUPLICATED on
every rank that touches them ("interface nodes", owner = lowest rank).  Per-cell work needs no
communication at all; the operator result needs one sum over the copies of each interface node per
apply (halo sum) and dot products count every node once (owner weights) before a scalar allreduce.

Everything here is host/torch index plumbing done once at setup; it also runs on CPU tensors so the
N>1 logic is tested with world_size-2 gloo jobs (tests/test_partition.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from .mesh import TetMesh


@dataclass
class Partition:
    rank: int
    n_ranks: int
    cell_range: tuple                  # [c0, c1) in the global (Morton) cell order
    local_nodes: torch.Tensor          # (M_loc,) global node id of every local node, ascending
    cells_local: torch.Tensor          # (N_loc,4) connectivity in local node ids
    owner_w: torch.Tensor              # (M_loc,) float64: 1 where this rank owns the node, else 0
    peers: list = field(default_factory=list)        # neighbour ranks, ascending
    shared: list = field(default_factory=list)       # per neighbour: (k,) local node ids, same global order on both sides
    n_global_cells: int = 0
    n_global_nodes: int = 0
    touch: torch.Tensor = None         # (M_global,) bitmask of the ranks touching each global node
    bounds: list = None                # cell ranges of ALL ranks: rank r holds [bounds[r], bounds[r+1])

    @property
    def n_interface(self):
        return int(sum(int(s.numel()) for s in self.shared))

    def local_mesh(self, mesh: TetMesh) -> TetMesh:
        """The rank's sub-mesh: its cells, their nodes, and the boundary triangles ASSIGNED to it.
        Every boundary triangle goes to exactly one rank -- the lowest one that has all three of its
        nodes -- so surface loads are partial sums that the halo sum completes.  Dirichlet node sets
        are taken from the GLOBAL triangles (a rank can hold a constrained node without holding any
        triangle of that boundary) and attached as ``boundary_nodes[tag]`` (local ids)."""
        ln = self.local_nodes.cpu().numpy()
        c0, c1 = self.cell_range
        lut = np.full(mesh.n_nodes, -1, dtype=np.int64)
        lut[ln] = np.arange(ln.size)
        touch = self.touch.cpu().numpy()
        if mesh.tris.shape[0]:
            m = touch[mesh.tris[:, 0]] & touch[mesh.tris[:, 1]] & touch[mesh.tris[:, 2]]
            owner_bit = m & (-m)
            keep = owner_bit == (1 << self.rank)
            tl = lut[mesh.tris[keep]]
        else:
            keep = np.zeros(0, dtype=bool)
            tl = mesh.tris
        out = TetMesh(mesh.coords[ln], self.cells_local.cpu().numpy(), mesh.cell_tags[c0:c1], tl,
                      mesh.tri_tags[keep], mesh.names)
        out.boundary_nodes = {}
        for tag in np.unique(mesh.tri_tags):
            nodes = np.unique(mesh.tris[mesh.tri_tags == tag])
            loc = lut[nodes]
            out.boundary_nodes[int(tag)] = loc[loc >= 0]
        return out


def chunk_bounds(n_cells, n_ranks):
    return [(n_cells * r) // n_ranks for r in range(n_ranks + 1)]


def build_partition(cells, n_nodes, rank, n_ranks, device="cpu", bounds=None) -> Partition:
    """cells: (N,4) global connectivity (Morton ordered).  Deterministic: every rank computes the same
    sharing pattern from the same global mesh, so no communication is needed to agree on the plan.
    bounds: the ranks' cell ranges (n_ranks + 1 ascending offsets) when they are not the equal chunks of the order --
    the nested partition of a multigrid hierarchy cuts a COARSE level into equal chunks and gives every finer cell
    to the rank of its ancestor (multigrid.nested_bounds)."""
    dev = torch.device(device)
    cells = torch.as_tensor(cells, device=dev).long()
    N = int(cells.shape[0])
    b = chunk_bounds(N, n_ranks) if bounds is None else [int(x) for x in bounds]
    if len(b) != n_ranks + 1 or b[0] != 0 or b[-1] != N or any(b[i] > b[i + 1] for i in range(n_ranks)):
        raise ValueError("bounds must be n_ranks + 1 ascending offsets from 0 to the number of cells")
    # bitmask of ranks touching each node (n_ranks <= 62)
    touch = torch.zeros(n_nodes, dtype=torch.int64, device=dev)
    for r in range(n_ranks):
        nodes_r = torch.unique(cells[b[r]:b[r + 1]].reshape(-1))
        touch[nodes_r] |= (1 << r)
    c0, c1 = b[rank], b[rank + 1]
    local_nodes = torch.nonzero(touch & (1 << rank)).reshape(-1)          # ascending global ids
    lut = torch.full((n_nodes,), -1, dtype=torch.int64, device=dev)
    lut[local_nodes] = torch.arange(local_nodes.numel(), device=dev)
    cells_local = lut[cells[c0:c1]]
    mask_loc = touch[local_nodes]
    lowest = mask_loc & (-mask_loc)                                       # lowest set bit = owner
    owner_w = (lowest == (1 << rank)).to(torch.float64)
    peers, shared = [], []
    for r in range(n_ranks):
        if r == rank:
            continue
        sel = torch.nonzero(mask_loc & (1 << r)).reshape(-1)              # local ids, ascending global order
        if sel.numel():
            peers.append(r)
            shared.append(sel)
    return Partition(rank, n_ranks, (c0, c1), local_nodes, cells_local, owner_w, peers, shared, N, n_nodes, touch, b)


def halo_sum_reference(part: Partition, vec: torch.Tensor, group=None):
    """Sum the copies of every interface node across ranks with torch.distributed point-to-point
    calls (any backend).  vec: (M_loc, ncomp).  Reference implementation of what sic_halo_sum does
    with NCCL send/recv inside the Krylov loop; used by the gloo tests and for setup-time vectors."""
    import torch.distributed as dist
    if part.n_ranks == 1:
        return vec
    sends = [vec[idx].contiguous() for idx in part.shared]
    recvs = [torch.empty_like(s) for s in sends]
    ops = []
    for peer, s, r in zip(part.peers, sends, recvs):
        ops.append(dist.P2POp(dist.isend, s, peer, group=group))
        ops.append(dist.P2POp(dist.irecv, r, peer, group=group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for idx, r in zip(part.shared, recvs):
        vec[idx] += r
    return vec
