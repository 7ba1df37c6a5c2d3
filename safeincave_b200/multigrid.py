"""Nested red-refinement hierarchy and the transfer tables of the geometric multigrid preconditioner
(host side: torch/numpy index plumbing, once per mesh; the arithmetic is csrc/mg.cu).

The synthetic 10M-80M cell cavern meshes of BASELINE config 5 are regular refinements of a gmsh grid
(SURVEY 8d), so every coarser level exists by construction.  ``refine_hierarchy`` keeps them all, each in
its own Morton order, together with, per level l >= 1 relative to level l-1:

    parent_a, parent_b  (M_l,)      the coarse node(s) a node interpolates from (a == b: it IS that node)
    rst_ptr, rst_idx    CSR         for every coarse node the level-l nodes it restricts from (weight 1/2 each)
    children            (8, N_{l-1}) the cells of level l refining each coarse cell
    inject              (M_{l-1},)  level-l id of every coarse node (Dirichlet masks are injected through it)
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .mesh import TetMesh, morton_order, red_refine


@dataclass
class Transfer:
    parent_a: np.ndarray
    parent_b: np.ndarray
    rst_ptr: np.ndarray
    rst_idx: np.ndarray
    children: np.ndarray
    inject: np.ndarray


@dataclass
class Hierarchy:
    meshes: list = field(default_factory=list)        # level 0 = coarsest ... last = finest (each Morton-ordered)
    transfers: list = field(default_factory=list)     # transfers[l] relates level l to l-1 (None for l = 0)

    @property
    def n_levels(self):
        return len(self.meshes)

    @property
    def finest(self) -> TetMesh:
        return self.meshes[-1]


def _transfer(coarse: TetMesh, fine_raw: TetMesh, edges: np.ndarray, fine: TetMesh) -> Transfer:
    Mc, Mf, Nc = coarse.n_nodes, fine.n_nodes, coarse.n_cells
    pa_raw = np.concatenate([np.arange(Mc, dtype=np.int64), edges[:, 0]])
    pb_raw = np.concatenate([np.arange(Mc, dtype=np.int64), edges[:, 1]])
    nperm, cperm = fine.node_perm, fine.cell_perm        # new id i was raw id perm[i]
    pa, pb = pa_raw[nperm], pb_raw[nperm]
    inv_c = np.empty(8 * Nc, dtype=np.int64)
    inv_c[cperm] = np.arange(8 * Nc)
    children = np.ascontiguousarray(inv_c.reshape(Nc, 8).T)
    inv_n = np.empty(Mf, dtype=np.int64)
    inv_n[nperm] = np.arange(Mf)
    inject = inv_n[:Mc]
    rows = np.concatenate([pa, pb])
    idx = np.concatenate([np.arange(Mf), np.arange(Mf)])
    order = np.argsort(rows, kind="stable")
    ptr = np.zeros(Mc + 1, dtype=np.int64)
    ptr[1:] = np.cumsum(np.bincount(rows, minlength=Mc))
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    return Transfer(i32(pa), i32(pb), i32(ptr), i32(idx[order]), i32(children), i32(inject))


def refine_hierarchy(base: TetMesh, levels: int, device="cpu", nested=False) -> Hierarchy:
    """``levels`` regular refinements of ``base``; returns all levels (level 0 = ``base`` in Morton order).
    nested: every refined level keeps the eight children of a cell together, in the order of the parents (cell 8 c + j
    of level l is child j of cell c of level l-1) -- a hierarchical space-filling order in which an equal-chunk cell
    partition of ANY level induces contiguous chunks on every finer level (what the multi-GPU multigrid partitions
    by, ``nested_bounds``); nodes are Morton-ordered either way."""
    m0 = base if hasattr(base, "cell_perm") else morton_order(base, device=device)
    h = Hierarchy([m0], [None])
    h.nested = bool(nested)
    for _ in range(levels):
        coarse = h.meshes[-1]
        raw, edges = red_refine(coarse, device=device, return_edges=True)
        fine = morton_order(raw, device=device, keep_cell_order=nested)
        h.transfers.append(_transfer(coarse, raw, edges, fine))
        h.meshes.append(fine)
    return h


def distributed_from(hierarchy: Hierarchy, n_ranks: int, min_cells_per_rank: int = 200_000) -> int:
    """Index of the coarsest level a multi-GPU multigrid PARTITIONS (every level from it to the finest is partitioned,
    the ones below are replicated on every rank): the coarsest level >= 1 that still gives every rank
    ``min_cells_per_rank`` cells -- below that a level's kernels are launch-latency bound and an exchange per operator
    application costs more than computing the level redundantly.  Without a nested hierarchy only the finest level can
    be partitioned."""
    top = hierarchy.n_levels - 1
    if n_ranks <= 1 or not getattr(hierarchy, "nested", False):
        return top
    lc = top
    while lc - 1 >= 1 and hierarchy.meshes[lc - 1].n_cells >= min_cells_per_rank * n_ranks:
        lc -= 1
    return lc


def nested_bounds(hierarchy: Hierarchy, n_ranks: int, level: int, lc: int):
    """Cell ranges of the ranks on ``level`` >= lc: level lc is cut into equal chunks of its (Morton) order and every
    finer cell belongs to the rank of its ancestor on level lc (offsets scale by 8 per level)."""
    from .partition import chunk_bounds
    return [b * 8 ** (level - lc) for b in chunk_bounds(hierarchy.meshes[lc].n_cells, n_ranks)]


def _localise(t: Transfer, part, n_coarse_nodes: int) -> Transfer:
    """Transfer tables of the finest level for ONE rank of a cell partition: rows of the rank's local nodes / cells,
    coarse side still in global numbering (the coarse level is replicated); children of other ranks become -1."""
    ln = part.local_nodes.cpu().numpy()
    c0, c1 = part.cell_range
    pa, pb = t.parent_a[ln].astype(np.int64), t.parent_b[ln].astype(np.int64)
    m = ln.size
    rows = np.concatenate([pa, pb])
    idx = np.concatenate([np.arange(m), np.arange(m)])
    order = np.argsort(rows, kind="stable")
    ptr = np.zeros(n_coarse_nodes + 1, dtype=np.int64)
    ptr[1:] = np.cumsum(np.bincount(rows, minlength=n_coarse_nodes))
    ch = t.children.astype(np.int64)
    ch = np.where((ch >= c0) & (ch < c1), ch - c0, -1)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    return Transfer(i32(pa), i32(pb), i32(ptr), i32(idx[order]), i32(ch), t.inject)


def _localise_pair(t: Transfer, part_f, part_c, n_coarse_nodes: int) -> Transfer:
    """Transfer tables between TWO partitioned levels of a nested partition, in the rank's local numbering on both
    sides: the parents of every local fine node are vertices of a local coarse cell, and the eight children of a local
    coarse cell are local fine cells."""
    ln_f, ln_c = part_f.local_nodes.cpu().numpy(), part_c.local_nodes.cpu().numpy()
    lut = np.full(n_coarse_nodes, -1, dtype=np.int64)
    lut[ln_c] = np.arange(ln_c.size)
    pa, pb = lut[t.parent_a[ln_f].astype(np.int64)], lut[t.parent_b[ln_f].astype(np.int64)]
    if (pa < 0).any() or (pb < 0).any():
        raise ValueError("the partitions of two multigrid levels are not nested (a fine node's parent is not local)")
    m, mc = ln_f.size, ln_c.size
    rows = np.concatenate([pa, pb])
    idx = np.concatenate([np.arange(m), np.arange(m)])
    order = np.argsort(rows, kind="stable")
    ptr = np.zeros(mc + 1, dtype=np.int64)
    ptr[1:] = np.cumsum(np.bincount(rows, minlength=mc))
    (f0, f1), (c0, c1) = part_f.cell_range, part_c.cell_range
    ch = t.children[:, c0:c1].astype(np.int64) - f0
    if ch.size and (ch.min() < 0 or ch.max() >= f1 - f0):
        raise ValueError("the partitions of two multigrid levels are not nested (a child cell is not local)")
    lut_f = np.full(t.parent_a.shape[0], -1, dtype=np.int64)
    lut_f[ln_f] = np.arange(m)
    inj = lut_f[t.inject[ln_c].astype(np.int64)]
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    return Transfer(i32(pa), i32(pb), i32(ptr), i32(idx[order]), i32(ch), i32(inj))


def prolongation_matrix(t: Transfer, n_coarse_nodes: int):
    """P (M_fine x M_coarse, scipy CSR) of one transfer: used by the tests to check the tables."""
    import scipy.sparse as sp
    Mf = t.parent_a.shape[0]
    rows = np.concatenate([np.arange(Mf), np.arange(Mf)])
    cols = np.concatenate([t.parent_a, t.parent_b]).astype(np.int64)
    return sp.csr_matrix((np.full(2 * Mf, 0.5), (rows, cols)), shape=(Mf, n_coarse_nodes))


# ----------------------------------------------------------------------------------------------
# device side: one operator-only Engine per coarse level + the sic_mg_level_t array
# ----------------------------------------------------------------------------------------------
class Multigrid:
    """V-cycle preconditioned CG on a ``Hierarchy`` (csrc/mg.cu through the C ABI).

    The finest level IS the momentum equation's engine (its C_T is the tangent of the Newton iteration);
    the coarse levels are operator-only engines whose C_T ``setup()`` fills by Galerkin coarsening."""

    def __init__(self, fine_engine, hierarchy: Hierarchy, nu=2, coarse_its=30, smooth_lo=0.1, coarse_lo=0.01,
                 safety=1.15, power_its=32, power_its_warm=4, part=None, coarse_fixed=None, dist=None,
                 dist_min_cells_per_rank=200_000, fused_coarse=True, use_graph=True, compressed=True):
        """part: the fine engine holds only this rank's cells (partition.Partition; ``hierarchy`` is the GLOBAL one).
        With a NESTED hierarchy (refine_hierarchy(..., nested=True)) and ``dist`` (the DistContext, for the levels'
        own exchange mailboxes) every level down to ``distributed_from(...)`` is partitioned by the rank's ancestors'
        cells; the levels below are replicated on every rank (csrc/mg.cu).  Otherwise only the finest level is.
        coarse_fixed: callable(level index, TetMesh) -> uint8 (3 M,) Dirichlet mask of a coarse level (global
        numbering); needed when the finest level is partitioned (a rank cannot inject a mask it holds a part of).
        fused_coarse: the coarsest level's sweep as one cooperative launch; use_graph: replay every Krylov iteration of
        ``solve`` from one captured CUDA graph (both: csrc/mg.cu; the graph is not used while the operator is being timed
        launch by launch).
        compressed: the operator applications INSIDE the V-cycle read a float copy of the symmetric part of C_T and of the
        geometry (152 B per cell instead of 408; sic_mg_level_t.pc_ct / pc_geom); the Krylov operator stays exact."""
        import ctypes

        import torch

        from . import _lib as L
        from .engine import _ptr
        self._L, self._ptr, self._ct, self._to = L, _ptr, ctypes, torch
        if hierarchy.n_levels > L.SIC_MG_MAX_LEVELS:
            raise L.SicError(f"at most {L.SIC_MG_MAX_LEVELS} multigrid levels")
        fine = hierarchy.finest
        self.part = part if (part is not None and part.n_ranks > 1) else None
        if self.part is None:
            if fine.n_cells != fine_engine.N or fine.n_nodes != fine_engine.M:
                raise L.SicError("the hierarchy's finest level is not the mesh of the momentum equation")
        else:
            c0, c1 = self.part.cell_range
            if c1 - c0 != fine_engine.N or int(self.part.local_nodes.numel()) != fine_engine.M:
                raise L.SicError("the partition does not describe the momentum equation's local mesh")
            if hierarchy.n_levels < 2 or coarse_fixed is None:
                raise L.SicError("a partitioned multigrid needs >= 2 levels and the coarse Dirichlet masks (coarse_fixed)")
            if fine_engine.halo is None:
                raise L.SicError("attach the partition to the engine (Engine.set_partition) before building the multigrid")
        self.coarse_fixed = coarse_fixed
        self.h, self.fine = hierarchy, fine_engine
        self.lib, dev = fine_engine.lib, fine_engine.device
        n = hierarchy.n_levels
        # per-level partitions of a nested multi-GPU hierarchy: parts[l] is None for a replicated level
        self.parts = [None] * n
        self.lc = n - 1
        if self.part is not None:
            self.parts[-1] = self.part
            lc = distributed_from(hierarchy, self.part.n_ranks, dist_min_cells_per_rank) if dist is not None else n - 1
            if lc < n - 1 and list(self.part.bounds) != nested_bounds(hierarchy, self.part.n_ranks, n - 1, lc):
                lc = n - 1            # the finest level was not partitioned by ancestors: replicate everything below it
            self.lc = lc
            from .partition import build_partition
            for l in range(lc, n - 1):
                m = hierarchy.meshes[l]
                self.parts[l] = build_partition(m.cells, m.n_nodes, self.part.rank, self.part.n_ranks, device=dev,
                                                bounds=nested_bounds(hierarchy, self.part.n_ranks, l, lc))
        self.engines = []
        for l, m in enumerate(hierarchy.meshes[:-1]):
            pl = self.parts[l]
            if pl is None:
                eng = type(fine_engine)(m.coords, m.cells, device=dev, operator_only=True)
            else:
                eng = type(fine_engine)(m.coords[pl.local_nodes.cpu().numpy()], pl.cells_local.cpu().numpy(), device=dev,
                                        operator_only=True)
                eng.set_partition(pl, dist.comm, dist.make_p2p(pl))
            self.engines.append(eng)
        self.engines.append(fine_engine)
        self.opts = L.SicMgOpts(int(nu), int(coarse_its), float(smooth_lo), float(coarse_lo), float(safety), int(power_its),
                                int(power_its_warm), 1 if fused_coarse else 0, 0)
        self.use_graph = bool(use_graph)
        self.graph_launches = 0
        self.levels = (L.SicMgLevel * n)()
        self._keep = []
        zeros = lambda *shape, dtype=torch.float64: torch.zeros(shape, dtype=dtype, device=dev)
        dt = lambda a: torch.as_tensor(a).to(dev).contiguous()
        self.fixed, self.dinv, self.vec = [], [], []
        for l, eng in enumerate(self.engines):
            self.fixed.append(zeros(3 * eng.M, dtype=torch.uint8))
            self.dinv.append(zeros(eng.M, 9))
            self.vec.append({k: zeros(3 * eng.M) for k in ("x", "b", "r", "d", "t", "pv")})
            lv = self.levels[l]
            lv.lambda_max = 0.0
            if l > 0:
                t = hierarchy.transfers[l]
                if self.parts[l] is not None and self.parts[l - 1] is not None:
                    t = _localise_pair(t, self.parts[l], self.parts[l - 1], hierarchy.meshes[l - 1].n_nodes)
                elif self.parts[l] is not None:
                    t = _localise(t, self.parts[l], hierarchy.meshes[l - 1].n_nodes)
                tabs = {k: dt(getattr(t, k)) for k in ("parent_a", "parent_b", "rst_ptr", "rst_idx", "children", "inject")}
                self._keep.append(tabs)
                lv.parent_a, lv.parent_b = _ptr(tabs["parent_a"]), _ptr(tabs["parent_b"])
                lv.rst_ptr, lv.rst_idx, lv.children = _ptr(tabs["rst_ptr"]), _ptr(tabs["rst_idx"]), _ptr(tabs["children"])
            else:
                self._keep.append(None)
            for k in ("x", "b", "r", "d", "t", "pv"):
                setattr(lv, k, _ptr(self.vec[l][k]))
            if compressed:
                pc_geom = torch.cat([eng.grad.to(torch.float32), eng.vol.to(torch.float32).reshape(1, -1)])
                pc_geom = pc_geom.reshape(13, eng.ns // 128, 128).permute(1, 0, 2).contiguous()      # [tile][13][128]
                pc_ct = torch.zeros((eng.ns // 128, 21, 128), dtype=torch.float32, device=dev)
                pc_dinv = torch.zeros(9 * max(eng.M, 1), dtype=torch.float32, device=dev)
                self.vec[l]["pc_geom"], self.vec[l]["pc_ct"], self.vec[l]["pc_dinv"] = pc_geom, pc_ct, pc_dinv
                lv.pc_ct, lv.pc_geom, lv.pc_lidx, lv.pc_dinv = _ptr(pc_ct), _ptr(pc_geom), _ptr(eng.lidx), _ptr(pc_dinv)
        self.compressed = bool(compressed)
        need = int(self.lib.sic_mg_workspace_doubles(fine_engine.N, fine_engine.M))
        self.work = zeros(need)
        # kernels per V-cycle (bench.py's gpu_launches): per level 2 nu operator + 2 nu smoother + residual, restriction,
        # prolongation; coarsest level 2 coarse_its - 1
        self.launches_per_cycle = (n - 1) * (4 * nu + 3) + 2 * coarse_its - 1
        self._coarse_launches = 2 * coarse_its - 1
        self.setups = 0

    def _refresh(self, fixed_fine=None, dinv_fine=None):
        """Re-point the level structs at the engines' CURRENT buffers (the fine engine swaps sig/eps)."""
        _ptr = self._ptr
        n = len(self.engines)
        if fixed_fine is not None:
            self.fixed[-1] = fixed_fine
        if dinv_fine is not None:
            self.dinv[-1] = dinv_fine
        for l, eng in enumerate(self.engines):
            lv = self.levels[l]
            lv.prob = eng.problem()
            lv.fixed, lv.dinv = _ptr(self.fixed[l]), _ptr(self.dinv[l])
        for l, eng in enumerate(self.engines):
            if self.parts[l] is not None:
                self.levels[l].halo = self._ct.cast(self._ct.pointer(eng.halo), self._ct.c_void_p)

    def setup(self, fixed_fine, dinv_fine):
        """Once per tangent: inject the Dirichlet mask down the hierarchy, then sic_mg_setup (Galerkin C_T,
        block-Jacobi blocks of every level, lambda_max)."""
        L, torch = self._L, self._to
        self._refresh(fixed_fine, dinv_fine)
        if self.part is None:
            for l in range(len(self.engines) - 1, 0, -1):
                inj = self._keep[l]["inject"].long()
                self.fixed[l - 1].view(-1, 3).copy_(self.fixed[l].view(-1, 3)[inj])
        else:       # every rank builds the coarse masks from the boundary data of the (global) coarse meshes, once per
            # set of Dirichlet boundaries; a partitioned coarse level keeps the rows of its local nodes
            bc = getattr(getattr(self.coarse_fixed, "__self__", None), "bc", None)
            key = tuple((b.boundary_name, int(b.component)) for b in bc.dirichlet_boundaries) if bc is not None else None
            if key is None or key != getattr(self, "_mask_key", ()):
                for l in range(len(self.engines) - 1):
                    m = np.asarray(self.coarse_fixed(l, self.h.meshes[l])).reshape(-1, 3)
                    if self.parts[l] is not None:
                        m = m[self.parts[l].local_nodes.cpu().numpy()]
                    self.fixed[l].copy_(self._to.as_tensor(np.ascontiguousarray(m)).to(self.fixed[l].device,
                                                                                   dtype=self._to.uint8).reshape(-1))
                self._mask_key = key
        st = self.fine._stream()
        L.check(self.lib.sic_mg_setup(self.levels, len(self.engines), self._ct.byref(self.opts), self._ptr(self.work), st),
                "sic_mg_setup")
        self.setups += 1
        its = self.opts.power_its_warm if (self.setups > 1 and self.opts.power_its_warm > 0) else self.opts.power_its
        self.fine.launches += sum(1 + 2 + 3 * max(its, 0) for _ in self.engines)

    def lambda_max(self):
        return [float(lv.lambda_max) for lv in self.levels]

    def vcycle(self, r):
        """z = M^-1 r (one V-cycle); r, z: (3 M_fine,) device tensors."""
        L = self._L
        self._refresh()
        self.vec[-1]["b"].copy_(r)
        L.check(self.lib.sic_mg_vcycle(self.levels, len(self.engines), self._ct.byref(self.opts), self._ptr(self.work),
                                       self.fine._stream()), "sic_mg_vcycle")
        return self.vec[-1]["x"]

    def solve(self, b_ext, x, rtol=1e-10, atol=0.0, max_it=500, check_every=4, guess_nonzero=False, time_operator=False):
        L, ct = self._L, self._ct
        self._refresh()
        ksp = L.SicKsp()
        ksp.method, ksp.max_it, ksp.rtol, ksp.atol = 0, int(max_it), float(rtol), float(atol)
        ksp.check_every, ksp.use_graph = int(check_every), 1 if self.use_graph else 0
        ksp.guess_nonzero = 1 if guess_nonzero else 0
        ksp.time_operator = 1 if time_operator else 0
        fused0 = int(self.lib.sic_mg_fused_coarse_launches())
        L.check(self.lib.sic_mg_solve(self.levels, len(self.engines), ct.byref(self.opts), ct.byref(ksp), self._ptr(b_ext),
                                      self._ptr(x), self._ptr(self.work), self.fine._stream()), "sic_mg_solve")
        eng = self.fine
        if int(self.lib.sic_mg_fused_coarse_launches()) > fused0 and self._coarse_launches > 1:
            self.launches_per_cycle -= self._coarse_launches - 1       # the coarsest sweep is ONE cooperative launch
            self._coarse_launches = 1
        # host-side launches: the prologue (incl. the first V-cycle) and the iterations launched kernel by kernel, plus
        # ONE graph launch per replayed iteration (graph_kernel_nodes: the kernels inside those graphs)
        per_it = self.launches_per_cycle + 5
        first_direct = 0 if int(ksp.graph_launches) > 0 else 1       # the first cycle of the solve is a graph of its own
        eng.launches += 6 + per_it * (int(ksp.direct_iterations) + first_direct) + int(ksp.graph_launches)
        eng.graph_kernel_nodes = getattr(eng, "graph_kernel_nodes", 0) + per_it * int(ksp.graph_launches)
        self.graph_launches += int(ksp.graph_launches)
        eng.op_ms += float(ksp.op_ms)
        eng.op_samples += int(ksp.op_samples)
        nu, n = self.opts.nu, len(self.engines)
        # finest-level operator launches: 2 nu per V-cycle (one cycle per iteration + the first), timed by op_ms -- the
        # compressed kernel when the levels carry pc_ct --, and the exact Krylov operator once per iteration (op_dot_ms)
        eng.op_launches += (2 * nu if n > 1 else self.opts.coarse_its) * (int(ksp.iterations) + 1)
        eng.op_dot_ms = getattr(eng, "op_dot_ms", 0.0) + float(ksp.op_dot_ms)
        eng.op_dot_samples = getattr(eng, "op_dot_samples", 0) + int(ksp.op_dot_samples)
        eng.op_dot_launches = getattr(eng, "op_dot_launches", 0) + int(ksp.iterations)
        eng.xchg_ms = getattr(eng, "xchg_ms", 0.0) + float(ksp.xchg_ms)
        eng.xchg_samples = getattr(eng, "xchg_samples", 0) + int(ksp.xchg_samples)
        return ksp
