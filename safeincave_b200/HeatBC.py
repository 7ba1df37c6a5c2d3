"""Boundary conditions of the heat equation with the interface of the reference's safeincave/HeatBC.py
(``DirichletBC``, ``NeumannBC``, ``RobinBC``, ``BcHandler``; :28-334).  Values are piecewise-linear in time
(``np.interp``, :261, :299, :331).  ``update_*`` fill per-triangle / per-node device arrays that csrc/heat.cu reads."""
from __future__ import annotations

from abc import ABC

import numpy as np
import torch as to


class GeneralBC(ABC):
    def __init__(self, boundary_name: str, values: list, time_values: list):
        self.boundary_name = boundary_name
        self.values = values
        self.time_values = time_values
        self.type = None


class DirichletBC(GeneralBC):
    def __init__(self, boundary_name: str, values: list, time_values: list):
        super().__init__(boundary_name, values, time_values)
        self.type = "dirichlet"


class NeumannBC(GeneralBC):
    def __init__(self, boundary_name: str, values: list, time_values: list):
        super().__init__(boundary_name, values, time_values)
        self.type = "neumann"


class RobinBC(GeneralBC):
    def __init__(self, boundary_name: str, values: list, h: float, time_values: list):
        super().__init__(boundary_name, values, time_values)
        self.type = "robin"
        self.h = h


class BcHandler:
    def __init__(self, equation):
        self.eq = equation
        self.reset_boundary_conditions()

    def reset_boundary_conditions(self) -> None:
        self.dirichlet_boundaries, self.neumann_boundaries, self.robin_boundaries = [], [], []
        self._cache = None

    def add_boundary_condition(self, bc: GeneralBC) -> None:
        if bc.type == "dirichlet":
            self.dirichlet_boundaries.append(bc)
        elif bc.type == "neumann":
            self.neumann_boundaries.append(bc)
        elif bc.type == "robin":
            self.robin_boundaries.append(bc)
        else:
            raise Exception(f"Boundary type {bc.type} not supported.")
        self._cache = None

    def _tables(self):
        if self._cache is None:
            eq = self.eq
            tm, dev = eq.grid.tetmesh, eq.engine.device
            sel = lambda bc: to.as_tensor(np.nonzero(tm.tri_tags == eq.grid.get_boundary_tag(bc.boundary_name))[0], device=dev)
            def nodes(bc):
                tag = eq.grid.get_boundary_tag(bc.boundary_name)
                if getattr(tm, "boundary_nodes", None) is not None:   # partitioned mesh: node sets from the GLOBAL mesh
                    n = np.asarray(tm.boundary_nodes.get(int(tag), np.zeros(0, dtype=np.int64)))
                else:
                    n = np.unique(tm.tris[tm.tri_tags == tag])
                return to.as_tensor(n, dtype=to.int64, device=dev)
            self._cache = ([nodes(bc) for bc in self.dirichlet_boundaries], [sel(bc) for bc in self.neumann_boundaries],
                           [sel(bc) for bc in self.robin_boundaries])
        return self._cache

    def update_bcs(self, t: float) -> None:
        """HeatBC.py:227-245."""
        self.update_dirichlet(t)
        self.update_neumann(t)
        self.update_robin(t)

    def update_dirichlet(self, t: float) -> None:
        """HeatBC.py:247-281: prescribed temperature on the nodes of the tagged facets (later BCs override earlier)."""
        eq = self.eq
        dn, _, _ = self._tables()
        eq.fixed.zero_()
        self.dirichlet_values = []
        for bc, nodes in zip(self.dirichlet_boundaries, dn):
            value = float(np.interp(t, bc.time_values, bc.values))
            self.dirichlet_values.append(value)
            eq.fixed[nodes] = 1
            eq.T_prescribed[nodes] = value

    def _flux_terms(self, t):
        eq = self.eq
        _, ns, rs = self._tables()
        eq.tri_q.zero_()
        eq.tri_h.zero_()
        for bc, sel in zip(self.neumann_boundaries, ns):          # HeatBC.py:283-304
            eq.tri_q[sel] += float(np.interp(t, bc.time_values, bc.values))
        for bc, sel in zip(self.robin_boundaries, rs):            # HeatBC.py:306-334
            T_inf = float(np.interp(t, bc.time_values, bc.values))
            eq.tri_h[sel] += float(bc.h)
            eq.tri_q[sel] += float(bc.h) * T_inf

    def update_neumann(self, t: float) -> None:
        self._flux_terms(t)

    def update_robin(self, t: float) -> None:
        self._flux_terms(t)
