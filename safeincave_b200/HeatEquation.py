"""``HeatDiffusion`` with the method surface of the reference's safeincave/HeatEquation.py:34-343, on the device.

    heat_eq = HeatDiffusion(grid); heat_eq.set_solver(ksp); heat_eq.set_material(mat)   # mat.k, mat.density, mat.cp
    heat_eq.set_initial_T(T_nodes); heat_eq.set_boundary_conditions(bc_handler)
    heat_eq.solve(t, dt); T_cells = heat_eq.get_T_elems()

The backward-Euler step (do.fem.form + assemble_matrix + assemble_vector + lifting + KSP.solve in the reference,
:304-343) is one call of ``sic_heat_step`` (csrc/heat.cu: matrix-free, Jacobi-preconditioned CG).  ``T`` / ``T_old`` /
``X`` keep the ``.x.array`` access of dolfinx Functions (host copies); ``get_T_elems`` returns a DEVICE tensor that
``LinearMomentum.set_T`` takes without a round trip through the host."""
from __future__ import annotations

import ctypes

import numpy as np
import torch as to

from . import _lib as L
from .engine import Engine, _ptr
from .MomentumEquation import _FunctionView
from .mesh import tri_area_normals


class HeatDiffusion:
    engine_cls = Engine      # the device back end (the test-suite's host emulation substitutes its own subclass)

    def __init__(self, grid, device="cuda"):
        self.grid = grid
        tm = grid.tetmesh
        self.engine = self.engine_cls(tm.coords, tm.cells, device=device, geometry_only=True)
        eng = self.engine
        # several GPUs (the reference's heat solve is MPI-parallel, HeatEquation.py:344-364): the grid of
        # distributed.partition_grid carries its cell partition and the process group; the heat equation gets its own
        # halo plan / exchange mailbox on the same partition (csrc/heat.cu)
        part = getattr(grid, "partition", None)
        self.dist = getattr(grid, "dist", None)
        if part is not None and part.n_ranks > 1:
            if self.dist is None:
                raise ValueError("a partitioned grid needs its DistContext (grid.dist, set by distributed.partition_grid)")
            eng.set_partition(part, self.dist.comm, self.dist.make_p2p(part))
        dev = eng.device
        self.n_elems, self.n_nodes = eng.N, eng.M
        z = lambda n, dt=to.float64: to.zeros(max(n, 1), dtype=dt, device=dev)
        self.T_dev, self.T_old_dev = z(eng.M), z(eng.M)
        self.T_prescribed = z(eng.M)
        self.fixed = z(eng.M, to.uint8)
        self.T_cells = z(eng.ns)
        self.rho_cp, self.k_dev = z(eng.ns), z(eng.ns)
        F = tm.tris.shape[0]
        self.n_tri = F
        self.tri = to.as_tensor(np.ascontiguousarray(tm.tris.T), dtype=to.int32, device=dev).contiguous() if F else z(1, to.int32)
        area = np.linalg.norm(tri_area_normals(tm), axis=1) if F else np.zeros(1)
        self.tri_area = to.as_tensor(area, dtype=to.float64, device=dev)
        self.tri_h, self.tri_q = z(F), z(F)
        self.work = z(int(eng.lib.sic_heat_workspace_doubles(eng.M)))
        self.solver, self.bc, self.mat = None, None, None
        self.ksp_log = []
        self.T = _FunctionView("T", lambda: self.T_dev[:self.n_nodes].cpu().numpy())
        self.T_old = _FunctionView("T_old", lambda: self.T_old_dev[:self.n_nodes].cpu().numpy())
        self.X = self.T

    # ------------------------------------------------------------------ configuration (HeatEquation.py:102-151)
    def set_material(self, material) -> None:
        self.mat = material
        self.initialize()

    def set_solver(self, solver) -> None:
        self.solver = solver

    def set_boundary_conditions(self, bc) -> None:
        self.bc = bc

    def initialize(self) -> None:
        """HeatEquation.py:219-234: k, rho, cp of the material become per-cell device arrays."""
        eng = self.engine
        for name in ("k", "density", "cp"):
            if not hasattr(self.mat, name):
                raise AttributeError(f"Material has no '{name}': call set_thermal_conductivity / set_density / "
                                     "set_specific_heat_capacity before HeatDiffusion.set_material")
        d = lambda v: to.as_tensor(v).to(eng.device, dtype=to.float64)
        self.k_dev[:eng.N] = d(self.mat.k)
        self.rho_cp[:eng.N] = d(self.mat.density) * d(self.mat.cp)

    def set_initial_T(self, T_field) -> None:
        """HeatEquation.py:266-284: nodal temperatures, in the order of ``grid.mesh.geometry.x``."""
        T = to.as_tensor(T_field).to(self.engine.device, dtype=to.float64).reshape(-1)
        self.T_dev[:self.n_nodes] = T
        self.T_old_dev[:self.n_nodes] = T

    def update_T_old(self) -> None:
        self.T_old_dev.copy_(self.T_dev)

    def split_solution(self) -> None:
        pass

    # ------------------------------------------------------------------ device calls
    def _problem(self):
        eng = self.engine
        H = L.SicHeat()
        H.n_cells, H.cell_stride, H.n_nodes, H.n_tri = eng.N, eng.ns, eng.M, self.n_tri
        H.conn, H.grad, H.vol = _ptr(eng.conn), _ptr(eng.grad), _ptr(eng.vol)
        H.rho_cp, H.k = _ptr(self.rho_cp), _ptr(self.k_dev)
        H.tri, H.tri_area, H.tri_h, H.tri_q = _ptr(self.tri), _ptr(self.tri_area), _ptr(self.tri_h), _ptr(self.tri_q)
        H.fixed = _ptr(self.fixed)
        H.halo = ctypes.cast(ctypes.pointer(eng.halo), ctypes.c_void_p) if eng.halo is not None else None
        return H

    def get_T_elems(self):
        """HeatEquation.py:286-302: the P1 field at the cells' interpolation point = mean of the four nodal values.
        Returns a (n_elems,) float64 tensor on the device."""
        eng = self.engine
        H = self._problem()
        L.check(eng.lib.sic_heat_cell_mean(ctypes.byref(H), _ptr(self.T_dev), _ptr(self.T_cells), eng._stream()),
                "sic_heat_cell_mean")
        eng.launches += 1
        return self.T_cells[:eng.N]

    def solve(self, t: float, dt: float) -> None:
        """HeatEquation.py:304-343: update the BCs, one backward-Euler step, T_old <- T."""
        eng, ksp = self.engine, self.solver
        if ksp is None or self.bc is None:
            raise RuntimeError("HeatDiffusion needs set_solver and set_boundary_conditions before solve")
        self.bc.update_bcs(t)
        T = self.T_dev
        to.where(self.fixed.bool(), self.T_prescribed, T, out=T)      # previous T is the initial guess
        rtol, atol, max_it = ksp.effective()
        res = L.SicKsp()
        res.method, res.max_it, res.rtol, res.atol, res.check_every = L.KSP_CG, int(min(max_it, 100000)), float(rtol), float(atol), 10
        H = self._problem()
        L.check(eng.lib.sic_heat_step(ctypes.byref(H), float(dt), _ptr(self.T_old_dev), _ptr(T), ctypes.byref(res),
                                      _ptr(self.work), eng._stream()), "sic_heat_step")
        eng.launches += (12 + 5 * int(res.iterations)) if eng.halo is None else (22 + 10 * int(res.iterations))
        ksp.record(res)
        self.ksp_log.append((int(res.iterations), int(res.reason), float(res.rnorm)))
        if res.reason < 0:
            raise L.SicError(f"heat solve did not converge (reason {res.reason} after {res.iterations} iterations)")
        self.update_T_old()
