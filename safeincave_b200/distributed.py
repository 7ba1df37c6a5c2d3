"""One process per GPU: torch.distributed for the plumbing, an NCCL communicator of our own for the
halo sums and scalar all-reduces issued from inside the CUDA library (include/safeincave_cuda.h,
``sic_comm_*``, ``sic_halo_sum``).  Launch with torchrun; RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*
come from the environment.
"""
from __future__ import annotations

import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib as L
from .Grid import GridHandlerGMSH
from .partition import Partition, build_partition


class DistContext:
    def __init__(self, rank, world, device, comm=None):
        self.rank, self.world, self.device, self.comm = rank, world, device, comm
        self.p2p = None
        self.use_p2p = os.environ.get("SIC_P2P", "1") != "0"    # 0: NCCL send/recv + allreduce instead

    def make_p2p(self, part):
        """Mailbox for the peer-to-peer exchange kernel: allocate, all-gather the IPC handles, map the peers."""
        if self.world == 1 or self.device.type != "cuda" or not self.use_p2p:
            return None
        lib = L.load()
        cap = max([int(s.numel()) for s in part.shared] + [1])
        t = torch.tensor([cap], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        handle = (ctypes.c_uint8 * 64)()
        ctx = ctypes.c_void_p(0)
        L.check(lib.sic_p2p_create(self.rank, self.world, int(t.item()), ctypes.byref(ctx),
                                   ctypes.cast(handle, ctypes.c_void_p)), "sic_p2p_create")
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
        allh = [torch.zeros(64, dtype=torch.uint8, device=self.device) for _ in range(self.world)]
        dist.all_gather(allh, mine)
        raw = bytes(torch.cat(allh).cpu().tolist())
        buf = (ctypes.c_uint8 * len(raw)).from_buffer_copy(raw)
        L.check(lib.sic_p2p_connect(ctx, ctypes.cast(buf, ctypes.c_void_p)), "sic_p2p_connect")
        dist.barrier()
        self.p2p = ctx
        return ctx

    def all_reduce_sum(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def barrier(self):
        if self.world > 1:
            dist.barrier()

    def max_over_ranks(self, value: float) -> float:
        if self.world == 1:
            return value
        t = torch.tensor([value], dtype=torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


def init(device=None) -> DistContext:
    """Initialise torch.distributed (NCCL) from the torchrun environment and create the library's own
    NCCL communicator.  With WORLD_SIZE unset or 1 nothing is initialised."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if device is None:
        device = torch.device("cuda", local)
    if device.type == "cuda":
        torch.cuda.set_device(device)
    if world == 1:
        return DistContext(0, 1, device, None)
    if not dist.is_initialized():
        dist.init_process_group(backend="nccl" if device.type == "cuda" else "gloo",
                                device_id=device if device.type == "cuda" else None)
    comm = None
    if device.type == "cuda":
        lib = L.load()
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (ctypes.c_uint8 * 128)()
            L.check(lib.sic_comm_unique_id(ctypes.cast(buf, ctypes.c_void_p)), "sic_comm_unique_id")
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.to(device)
        dist.broadcast(uid, src=0)
        raw = bytes(uid.cpu().tolist())
        cbuf = (ctypes.c_uint8 * 128).from_buffer_copy(raw)
        handle = ctypes.c_void_p(0)
        L.check(lib.sic_comm_init(ctypes.cast(cbuf, ctypes.c_void_p), rank, world, ctypes.byref(handle)), "sic_comm_init")
        comm = handle
    return DistContext(rank, world, device, comm)


def partition_grid(ctx: DistContext, tetmesh, hierarchy=None, min_cells_per_rank=200_000):
    """Every rank holds the same global (Morton-ordered) mesh; returns (local grid, Partition).
    hierarchy: the GLOBAL multigrid hierarchy whose finest level is ``tetmesh`` (kept on the local grid for PC mg).
    With a nested hierarchy (multigrid.refine_hierarchy(..., nested=True)) the cells are cut by their ancestors on the
    coarsest level that still gives every rank ``min_cells_per_rank`` cells, so that the multigrid can partition every
    level from there up (multigrid.distributed_from); otherwise into equal chunks of the finest level's order."""
    bounds = None
    if hierarchy is not None and getattr(hierarchy, "nested", False) and ctx.world > 1:
        from .multigrid import distributed_from, nested_bounds
        lc = distributed_from(hierarchy, ctx.world, min_cells_per_rank)
        bounds = nested_bounds(hierarchy, ctx.world, hierarchy.n_levels - 1, lc)
    part = build_partition(tetmesh.cells, tetmesh.n_nodes, ctx.rank, ctx.world,
                           device=ctx.device if ctx.device.type == "cuda" else "cpu", bounds=bounds)
    local = part.local_mesh(tetmesh)
    grid = GridHandlerGMSH.from_mesh(local, reorder=False)
    grid.partition = part
    grid.dist = ctx
    grid.dist_min_cells_per_rank = min_cells_per_rank
    if hierarchy is not None:
        grid.hierarchy = hierarchy
    return grid, part


def attach(eq, part: Partition, ctx: DistContext):
    """Give a LinearMomentum built on the local grid its halo plan and communicator."""
    eq.dist = ctx
    if ctx.world > 1:
        eq.engine.set_partition(part, ctx.comm, ctx.make_p2p(part))
