"""CPU oracle of the heat equation and of the thermo-mechanical time loop.

TEST INFRASTRUCTURE ONLY (tests/ import it; the product never does).

PARITY UNPINNED BY THE REFERENCE, like oracle/fem.py: the arithmetic of safeincave/HeatEquation.py is done by
DOLFINx/FFCx/PETSc (absent, not installable) and no reference test touches it.  This module restates what the
UFL forms MEAN on P1 tetrahedra, following the call sites:

    HeatEquation.py:304-343   a = (rho cp dT T_/dt + k grad dT . grad T_) dx + sum_robin h dT T_ ds
                              L = (rho cp T_old T_/dt) dx + sum_neumann q T_ ds + sum_robin h T_inf T_ ds
                              Dirichlet: assemble_matrix(bcs) + apply_lifting + set_bc; backward Euler; T_old <- T
    HeatBC.py:247-334         values = np.interp(t, time_values, values); Dirichlet dofs = nodes of the tagged facets
    HeatEquation.py:286-302   get_T_elems: interpolation of the P1 field at the cell's interpolation point = mean of
                              its four nodal values
    Simulators.py:92-270      Simulator_TM.run: heat step, set_T, then the fixed-point loop of the momentum step with
                              tol 1e-6 / maxiter 20, NO dt-retry and an unconditional commit

with exact element matrices (consistent mass V/20 (1 + delta_ab), boundary mass A/12 (1 + delta_ab)), an assembled
scipy CSR matrix and a sparse direct solve -- a different formulation from the matrix-free CUDA kernels.  Pinned by
self-evident properties in tests/test_oracle_heat.py (linear steady profile reproduced exactly, energy balance,
Robin relaxation to T_inf, second-order spatial convergence against the 1-D analytic solution).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import constitutive as oc
from . import fem


def mass_stiffness(coords, cells, rho_cp, k):
    """M = int rho cp phi_a phi_b dx, K = int k grad phi_a . grad phi_b dx (CSR, M_nodes x M_nodes)."""
    grad, vol = fem.tet_geometry(coords, cells)
    N, M = cells.shape[0], coords.shape[0]
    Me = (vol * rho_cp / 20.0)[:, None, None] * (np.ones((4, 4)) + np.eye(4))[None]
    Ke = (vol * k)[:, None, None] * np.einsum("nai,nbi->nab", grad, grad)
    rows = np.repeat(cells, 4, axis=1).ravel()
    cols = np.tile(cells, (1, 4)).ravel()
    Mm = sp.csr_matrix((Me.ravel(), (rows, cols)), shape=(M, M))
    Km = sp.csr_matrix((Ke.ravel(), (rows, cols)), shape=(M, M))
    return Mm, Km


def tri_areas(coords, tris):
    a, b, c = (coords[tris[:, i]] for i in range(3))
    return 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)


def boundary_terms(coords, tris, tri_tags, neumann, robin, t):
    """(R, q): R = sum_robin int h phi_a phi_b ds (CSR), q_a = sum_neumann int q phi_a ds + sum_robin int h T_inf phi_a ds."""
    M = coords.shape[0]
    R = sp.csr_matrix((M, M))
    q = np.zeros(M)
    area = tri_areas(coords, tris)
    for bc in neumann:
        sel = tri_tags == bc["tag"]
        val = np.interp(t, bc["time_values"], bc["values"])
        np.add.at(q, tris[sel].ravel(), np.repeat(val * area[sel] / 3.0, 3))
    for bc in robin:
        sel = tri_tags == bc["tag"]
        T_inf = np.interp(t, bc["time_values"], bc["values"])
        tr, ar = tris[sel], area[sel]
        Re = (bc["h"] * ar / 12.0)[:, None, None] * (np.ones((3, 3)) + np.eye(3))[None]
        rows = np.repeat(tr, 3, axis=1).ravel()
        cols = np.tile(tr, (1, 3)).ravel()
        R = R + sp.csr_matrix((Re.ravel(), (rows, cols)), shape=(M, M))
        np.add.at(q, tr.ravel(), np.repeat(bc["h"] * T_inf * ar / 3.0, 3))
    return R, q


def dirichlet_nodes(tris, tri_tags, dirichlet, t):
    nodes, vals = [], []
    for bc in dirichlet:
        n = np.unique(tris[tri_tags == bc["tag"]])
        nodes.append(n)
        vals.append(np.full(n.size, np.interp(t, bc["time_values"], bc["values"])))
    if not nodes:
        return np.zeros(0, dtype=np.int64), np.zeros(0)
    nodes, vals = np.concatenate(nodes), np.concatenate(vals)
    _, first = np.unique(nodes[::-1], return_index=True)      # later BCs override earlier ones (order of set_bc)
    keep = len(nodes) - 1 - first
    return nodes[keep], vals[keep]


class OracleHeat:
    def __init__(self, coords, cells, tris, tri_tags, rho, cp, k, dirichlet, neumann, robin):
        self.coords, self.cells, self.tris, self.tri_tags = coords, cells, tris, tri_tags
        self.M, self.K = mass_stiffness(coords, cells, np.asarray(rho, float) * np.asarray(cp, float), np.asarray(k, float))
        self.dirichlet, self.neumann, self.robin = dirichlet, neumann, robin
        self.T = np.zeros(coords.shape[0])
        self.T_old = np.zeros(coords.shape[0])

    def set_initial_T(self, T):
        self.T = np.asarray(T, float).copy()
        self.T_old = self.T.copy()

    def cell_mean(self):
        return self.T[self.cells].mean(axis=1)

    def step(self, t, dt):
        R, q = boundary_terms(self.coords, self.tris, self.tri_tags, self.neumann, self.robin, t)
        A = (self.M / dt + self.K + R).tocsr()
        b = (self.M / dt) @ self.T_old + q
        nodes, vals = dirichlet_nodes(self.tris, self.tri_tags, self.dirichlet, t)
        n = self.coords.shape[0]
        T = np.zeros(n)
        T[nodes] = vals
        free = np.ones(n, dtype=bool)
        free[nodes] = False
        Ac = A.tocsc()
        rhs = b[free] - Ac[free][:, ~free] @ T[~free]
        T[free] = spla.spsolve(Ac[free][:, free].tocsc(), rhs)
        self.T = T
        self.T_old = T.copy()
        return T


class OracleSimulatorTM:
    """Simulator_TM.run (Simulators.py:92-270) on top of OracleHeat, OracleMaterial and oracle/fem.py."""

    def __init__(self, mech: fem.OracleSimulatorM, heat: OracleHeat):
        self.mech, self.heat = mech, heat
        self.history = []

    def run(self, t0, dt_list):
        s, h, m = self.mech, self.heat, self.mech.mat
        th = s.theta
        t = t0
        T_el = h.cell_mean()
        s.T0 = T_el.copy()                                      # :137-138
        s.T = T_el.copy()
        if s.compute_elastic_response:                          # :144-153
            s.u = s._solve(m.C, np.zeros((m.n, 6)), t)
            eps = fem.strain(s.coords, s.cells, s.u)
            sig = m.elastic_stress(eps)
        else:
            eps = fem.strain(s.coords, s.cells, s.u)
            sig = s.sig.copy()
        m.eval_rates(sig, t * th, s.T)                          # :168 passes t as dt (T7)
        m.commit_rates()
        self.history.append(dict(t=t, u=s.u.copy(), sig=sig.copy(), T=h.T.copy(), iters=0))
        sig_k = sig.copy()
        for dt in dt_list:
            t = t + dt
            h.step(t, dt)                                       # :196
            s.T = h.cell_mean()                                 # :199-200
            tol, err, ite = 1e-6, 2e-6, 0
            while err > tol and ite < 20:
                eps_k, sig_k = eps.copy(), sig.copy()
                CT, eps_rhs = m.tangent_phase(sig_k, s.T, s.T0, dt, th)
                s.u = s._solve(CT, eps_rhs, t)
                eps = fem.strain(s.coords, s.cells, s.u)
                sig = m.post_phase(eps, sig_k, s.T, dt, th)
                err = 0.0 if (th == 1.0 or not m.elems) else oc.newton_error(eps_k, eps)
                ite += 1
            m.commit(sig, sig_k, dt, th)                        # unconditional (T14)
            s.sig = sig.copy()
            self.history.append(dict(t=t, u=s.u.copy(), sig=sig.copy(), eps=eps.copy(), T=h.T.copy(), iters=ite, error=err))
        return self.history
