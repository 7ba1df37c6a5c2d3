"""Boundary conditions with the interface of the reference's safeincave/MomentumBC.py
(``DirichletBC`` :52-83, ``NeumannBC`` :85-135, ``BcHandler`` :138-277).

``update_dirichlet(t)`` turns the time-interpolated values into a constrained-dof mask plus the
prescribed values on the device (what ``locate_dofs_topological`` + ``dirichletbc`` give the
reference, :231-245); ``update_neumann(t)`` evaluates the surface load
``int (p(t) + rho g (H - x_i)) n.v ds`` (:270-277, note p = -interp) with the ``sic_neumann``
kernel into the external-load vector.
"""
from __future__ import annotations

import numpy as np
import torch as to

from .mesh import tri_area_normals


class GeneralBC:
    def __init__(self):
        self.boundary_name = None
        self.type = None
        self.values = None
        self.time_values = None


class DirichletBC(GeneralBC):
    def __init__(self, boundary_name, component, values, time_values):
        self.boundary_name = boundary_name
        self.type = "dirichlet"
        self.values = values
        self.time_values = time_values
        self.component = component


class NeumannBC(GeneralBC):
    def __init__(self, boundary_name, direction, density, ref_pos, values, time_values, g=-9.81):
        self.boundary_name = boundary_name
        self.type = "neumann"
        self.values = values
        self.time_values = time_values
        self.direction = direction
        self.density = density
        self.ref_pos = ref_pos
        self.gravity = g


class BcHandler:
    def __init__(self, equation):
        self.eq = equation
        self.dirichlet_boundaries = []
        self.neumann_boundaries = []
        self._dev_cache = None

    def reset_boundary_conditions(self):
        self.dirichlet_boundaries = []
        self.neumann_boundaries = []
        self._dev_cache = None

    def add_boundary_condition(self, bc):
        if bc.type == "dirichlet":
            self.dirichlet_boundaries.append(bc)
        elif bc.type == "neumann":
            self.neumann_boundaries.append(bc)
        else:
            raise Exception(f"Boundary type {bc.type} not supported.")
        self._dev_cache = None

    # ------------------------------------------------------------------ device-side tables
    def _tables(self):
        if self._dev_cache is not None:
            return self._dev_cache
        grid, eng = self.eq.grid, self.eq.engine
        tm, dev = grid.tetmesh, eng.device
        dofs = []
        for bc in self.dirichlet_boundaries:
            tag = grid.get_boundary_tag(bc.boundary_name)
            if getattr(tm, "boundary_nodes", None) is not None:     # partitioned mesh: node sets from the global mesh
                nodes = np.asarray(tm.boundary_nodes.get(int(tag), np.zeros(0, dtype=np.int64)))
            else:
                nodes = np.unique(tm.tris[tm.tri_tags == tag])
            dofs.append(to.as_tensor(3 * nodes + int(bc.component), dtype=to.int64, device=dev))
        neu = []
        if self.neumann_boundaries:
            tri = to.as_tensor(np.ascontiguousarray(tm.tris.T), dtype=to.int32, device=dev).contiguous()
            area_n = to.as_tensor(np.ascontiguousarray(tri_area_normals(tm).T), dtype=to.float64, device=dev).contiguous()
            for bc in self.neumann_boundaries:
                tag = grid.get_boundary_tag(bc.boundary_name)
                sel = np.where(tm.tri_tags == tag, 0, -1).astype(np.int32)
                if getattr(tm, "tri_interior", None) is not None and tm.tri_interior[sel == 0].any():
                    raise ValueError(f"Neumann boundary '{bc.boundary_name}' contains interior facets")
                neu.append(to.as_tensor(sel, device=dev))
            self._tri, self._area_n = tri, area_n
        self._dev_cache = (dofs, neu)
        return self._dev_cache

    def update_dirichlet(self, t):
        eq = self.eq
        dofs, _ = self._tables()
        eq.fixed.zero_()
        self.dirichlet_values = []
        for bc, d in zip(self.dirichlet_boundaries, dofs):
            value = float(np.interp(t, bc.time_values, bc.values))
            self.dirichlet_values.append(value)
            eq.fixed[d] = 1
            eq.u_prescribed[d] = value
        self.dirichlet_bcs = list(zip(self.dirichlet_boundaries, self.dirichlet_values))

    def update_neumann(self, t):
        eq = self.eq
        _, neu = self._tables()
        eq.b_neumann.zero_()
        self.neumann_bcs = []
        for bc, sel in zip(self.neumann_boundaries, neu):
            p = -float(np.interp(t, bc.time_values, bc.values))
            par = to.tensor([[p, bc.density * bc.gravity, bc.ref_pos, float(bc.direction)]], dtype=to.float64,
                            device=eq.engine.device)
            if self._tri.shape[1] > 0:
                eq.engine.neumann(self._tri, self._area_n, sel, par, eq.b_neumann)
            self.neumann_bcs.append((bc, p))
        eq.engine.halo_sum(eq.b_neumann, 3)          # several GPUs: complete the interface nodes
        to.add(eq.b_body, eq.b_neumann, out=eq.b_ext)
