"""CPU oracle for hot-path parts (2) assembly and (3) solve, and for the time loop around them.

TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py; never by the product).

PARITY UNPINNED BY THE REFERENCE: in the reference this arithmetic is done by third-party native
code that is absent from /root/reference and cannot be installed here -- DOLFINx 0.9 + FFCx/Basix/
UFL (assembly, README.md:5), PETSc via petsc4py (KSP), mpi4py, meshio/gmsh -- and no
reference test touches assembly, boundary conditions or the solve.  This module restates what the
UFL forms of the reference MEAN on P1 tetrahedra, following the call sites:

    MomentumEquation.py:1008-1011  a(u,v) = int (C_T : eps(u)) : eps(v) dx       -> assemble_K
    MomentumEquation.py:1014-1016  L(v) = body + Neumann + int (C_T:eps_rhs):eps(v) dx -> rhs_*
    MomentumEquation.py:1010,1017-1020 / MomentumBC.py:211-245  Dirichlet elimination -> solve
    MomentumBC.py:247-277          Neumann (p(t) + rho g (H - x_i)) n.v ds        -> neumann_load
    Utils.py:83-136, MomentumEquation.py:326-341  eps = sym grad u at DG0         -> strain
    Grid.py:139-242                volumes, node<->cell smoother                  -> smoother
    Simulators.py:310-541          Simulator_M time / Newton / dt-retry loop      -> OracleSimulatorM

and is pinned by self-evident properties in tests/test_oracle_fem.py (patch test, rigid-body null
space, symmetry for an elastic tangent, total Neumann force, manufactured solution); its outputs
are then the goldens for the CUDA path.  It deliberately uses a DIFFERENT formulation from the
CUDA kernels: explicit 6x12 B matrices and an assembled scipy CSR matrix with a sparse direct
solve, versus the kernels' matrix-free stress form and Krylov iteration.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import constitutive as oc

W_VOIGT = np.array([1.0, 1.0, 1.0, 2.0, 2.0, 2.0])


def tet_geometry(coords, cells):
    """grad phi_a (N,4,3) and volume (N,) of P1 tets (volume formula of Grid.py:115-137)."""
    x = coords[cells]                                   # (N,4,3)
    J = np.stack([x[:, 1] - x[:, 0], x[:, 2] - x[:, 0], x[:, 3] - x[:, 0]], axis=2)   # columns = edges
    det = np.linalg.det(J)
    Jinv = np.linalg.inv(J)                             # rows = grad lambda_1..3
    g123 = Jinv
    g0 = -g123.sum(axis=1, keepdims=True)
    return np.concatenate([g0, g123], axis=1), np.abs(det) / 6.0


def B_matrices(grad):
    """Tensorial strain-displacement matrices (N,6,12): eps_v = B u_e with
    eps_v = [xx,yy,zz,xy,xz,yz] (no factor 2, Utils.py:196) and u_e = [u0x,u0y,u0z,u1x,...]."""
    N = grad.shape[0]
    B = np.zeros((N, 6, 12))
    for a in range(4):
        gx, gy, gz = grad[:, a, 0], grad[:, a, 1], grad[:, a, 2]
        B[:, 0, 3 * a + 0] = gx
        B[:, 1, 3 * a + 1] = gy
        B[:, 2, 3 * a + 2] = gz
        B[:, 3, 3 * a + 0] = 0.5 * gy
        B[:, 3, 3 * a + 1] = 0.5 * gx
        B[:, 4, 3 * a + 0] = 0.5 * gz
        B[:, 4, 3 * a + 2] = 0.5 * gx
        B[:, 5, 3 * a + 1] = 0.5 * gz
        B[:, 5, 3 * a + 2] = 0.5 * gy
    return B


def cell_dofs(cells):
    return (3 * cells[:, :, None] + np.arange(3)[None, None, :]).reshape(cells.shape[0], 12)


def assemble_K(coords, cells, CT):
    """K = sum_e V_e B^T W C_T B as CSR (3M x 3M); C_T used as stored (not symmetrised)."""
    grad, vol = tet_geometry(coords, cells)
    B = B_matrices(grad)
    WC = W_VOIGT[None, :, None] * CT                                    # (N,6,6)
    Ke = vol[:, None, None] * np.matmul(B.transpose(0, 2, 1), np.matmul(WC, B))   # (N,12,12): V B^T (W C_T B)
    dofs = cell_dofs(cells)
    rows = np.repeat(dofs, 12, axis=1).ravel()
    cols = np.tile(dofs, (1, 12)).ravel()
    M3 = 3 * coords.shape[0]
    return sp.csr_matrix((Ke.ravel(), (rows, cols)), shape=(M3, M3))


def rhs_eps(coords, cells, CT, eps_rhs):
    """int (C_T : eps_rhs) : eps(v) dx  ->  sum_e V_e B^T W (C_T eps_rhs)."""
    grad, vol = tet_geometry(coords, cells)
    B = B_matrices(grad)
    s = oc.ddot(CT, eps_rhs) * W_VOIGT[None, :]
    fe = vol[:, None] * np.einsum("nia,ni->na", B, s)
    b = np.zeros(3 * coords.shape[0])
    np.add.at(b, cell_dofs(cells).ravel(), fe.ravel())
    return b


def body_force(coords, cells, density, g):
    """int rho g . v dx (MomentumEquation.py:255-275): rho_e g V_e / 4 on each node of the cell."""
    _, vol = tet_geometry(coords, cells)
    b = np.zeros((coords.shape[0], 3))
    f = (density * vol / 4.0)[:, None] * np.asarray(g, dtype=np.float64)[None, :]
    for a in range(4):
        np.add.at(b, cells[:, a], f)
    return b.ravel()


def neumann_load(coords, tris, tri_tags, bcs, t):
    """sum over Neumann BCs of int_F (p + rho g (H - x_i)) n.v ds, p = -interp(t) (MomentumBC.py:270-277).
    tris are outward oriented.  bcs: iterable of dicts(tag, direction, density, ref_pos, gravity,
    values, time_values).  The integrand is linear on the facet -> 3-point (edge-midpoint) rule is exact."""
    b = np.zeros((coords.shape[0], 3))
    for bc in bcs:
        sel = tris[tri_tags == bc["tag"]]
        if sel.size == 0:
            continue
        p = -np.interp(t, bc["time_values"], bc["values"])
        xa, xb, xc = (coords[sel[:, k]] for k in range(3))
        an = 0.5 * np.cross(xb - xa, xc - xa)           # area * outward normal
        f = lambda x: p + bc["density"] * bc["gravity"] * (bc["ref_pos"] - x[:, bc["direction"]])
        mids = [(0.5 * (xa + xb), (0.5, 0.5, 0.0)), (0.5 * (xb + xc), (0.0, 0.5, 0.5)), (0.5 * (xa + xc), (0.5, 0.0, 0.5))]
        for k in range(3):
            w = sum(f(xm) * phi[k] for xm, phi in mids) / 3.0
            np.add.at(b, sel[:, k], w[:, None] * an)
    return b.ravel()


def dirichlet_dofs(tris, tri_tags, tag, component):
    """dofs = component c of every node on facets carrying the tag (MomentumBC.py:231-245)."""
    nodes = np.unique(tris[tri_tags == tag])
    return 3 * nodes + component


def solve(K, b, fixed_dofs, fixed_vals):
    """K_ff u_f = b_f - K_fc u_c, u_c prescribed: what assemble_matrix(bcs) + apply_lifting +
    set_bc + KSP.solve compute (MomentumEquation.py:1010-1025), with an exact sparse LU."""
    n = K.shape[0]
    u = np.zeros(n)
    fixed_dofs = np.asarray(fixed_dofs, dtype=np.int64)
    u[fixed_dofs] = fixed_vals
    free = np.ones(n, dtype=bool)
    free[fixed_dofs] = False
    Kc = K.tocsc()
    rhs = b[free] - Kc[free][:, ~free] @ u[~free]
    u[free] = spla.spsolve(Kc[free][:, free].tocsc(), rhs)
    return u


def strain(coords, cells, u):
    """eps = sym grad u per cell, Voigt (N,6) (compute_total_strain, MomentumEquation.py:326-341)."""
    grad, _ = tet_geometry(coords, cells)
    B = B_matrices(grad)
    ue = u.reshape(-1, 3)[cells].reshape(cells.shape[0], 12)
    return np.einsum("nia,na->ni", B, ue)


def smoothers(coords, cells):
    """A (nodes x cells, volume weighted) and B (cells x nodes, 1/4) of Grid.py:198-242."""
    _, vol = tet_geometry(coords, cells)
    N, M = cells.shape[0], coords.shape[0]
    rows = cells.ravel()
    cols = np.repeat(np.arange(N), 4)
    A = sp.csr_matrix((np.repeat(vol, 4), (rows, cols)), shape=(M, N))
    A = sp.diags(1.0 / np.asarray(A.sum(axis=1)).ravel()) @ A
    Bm = sp.csr_matrix((np.full(4 * N, 0.25), (cols, rows)), shape=(N, M))
    return A, Bm


def p_q_fields(coords, cells, sig):
    """compute_p_elems/q_elems/p_nodes/q_nodes (MomentumEquation.py:287-324, 944-976)."""
    A, Bm = smoothers(coords, cells)
    I1 = sig[:, 0] + sig[:, 1] + sig[:, 2]
    I2 = sig[:, 0] * sig[:, 1] + sig[:, 1] * sig[:, 2] + sig[:, 0] * sig[:, 2] - sig[:, 3] ** 2 - sig[:, 4] ** 2 - sig[:, 5] ** 2
    p = I1 / 3.0
    q = np.sqrt(3 * ((1 / 3) * I1 ** 2 - I2))
    return dict(p_nodes=A @ p, q_nodes=A @ q, p_elems=Bm @ (A @ p), q_elems=Bm @ (A @ q))


class OracleSimulatorM:
    """Restatement of Simulator_M.run (Simulators.py:310-541) on top of OracleMaterial and the
    assembled operator above; tol 1e-8, maxiter 40, <= 3 dt halvings (T13 reproduced: a retry keeps
    t and the BC values of the full step)."""

    def __init__(self, coords, cells, tris, tri_tags, mat: oc.OracleMaterial, theta, T, T0, density, g,
                 dirichlet, neumann, compute_elastic_response=True):
        self.coords, self.cells, self.tris, self.tri_tags = coords, cells, tris, tri_tags
        self.mat, self.theta, self.T, self.T0 = mat, theta, np.asarray(T, float), np.asarray(T0, float)
        self.b_body = body_force(coords, cells, np.asarray(density, float), g)
        self.dirichlet, self.neumann = dirichlet, neumann     # lists of dicts
        self.compute_elastic_response = compute_elastic_response
        self.u = np.zeros(3 * coords.shape[0])
        self.sig = np.zeros((cells.shape[0], 6))
        self.history = []
        self.tol, self.maxiter, self.max_dt_cuts = 1e-8, 40, 3     # Simulators.py:399-402, 378-396
        self.after_initial_stress = None      # callable(material, sig) between initial stress and initial rates

    def _bc(self, t):
        dofs, vals = [], []
        for bc in self.dirichlet:
            d = dirichlet_dofs(self.tris, self.tri_tags, bc["tag"], bc["component"])
            dofs.append(d)
            vals.append(np.full(d.size, np.interp(t, bc["time_values"], bc["values"])))
        dofs = np.concatenate(dofs) if dofs else np.zeros(0, dtype=np.int64)
        vals = np.concatenate(vals) if vals else np.zeros(0)
        # later BCs override earlier ones on shared dofs (order of set_bc)
        _, first = np.unique(dofs[::-1], return_index=True)
        keep = len(dofs) - 1 - first
        b_ext = self.b_body + neumann_load(self.coords, self.tris, self.tri_tags, self.neumann, t)
        return dofs[keep], vals[keep], b_ext

    def _solve(self, CT, eps_rhs, t):
        dofs, vals, b_ext = self._bc(t)
        K = assemble_K(self.coords, self.cells, CT)
        b = b_ext + rhs_eps(self.coords, self.cells, CT, eps_rhs)
        return solve(K, b, dofs, vals)

    def run(self, t0, dt_list, maxiter_list=None):
        """maxiter_list: optional Newton-iteration cap per step (tests of the dt-retry path); default self.maxiter."""
        m, th = self.mat, self.theta
        t = t0
        if self.compute_elastic_response:                      # Simulators.py:346-354
            self.u = self._solve(m.C, np.zeros((m.n, 6)), t)
            eps = strain(self.coords, self.cells, self.u)
            sig = m.elastic_stress(eps)
        else:
            eps = strain(self.coords, self.cells, self.u)
            sig = self.sig.copy()
        if self.after_initial_stress is not None:
            self.after_initial_stress(m, sig)
        m.eval_rates(sig, t * th, self.T)                      # :364 passes t as dt (T7)
        m.commit_rates()                                       # :365
        self.history.append(dict(t=t, u=self.u.copy(), sig=sig.copy(), eps=eps.copy(), iters=0, error=0.0))
        for k_step, dt in enumerate(dt_list):
            t = t + dt
            if maxiter_list is not None:
                self.maxiter = maxiter_list[k_step]
            sig_bak, eps_bak, snap = sig.copy(), eps.copy(), m.snapshot()
            dt_cur, cuts, converged = dt, 0, False
            while not converged and cuts <= self.max_dt_cuts:
                tol, err, ite = self.tol, 2 * self.tol, 0
                while err > tol and ite < self.maxiter:
                    eps_k, sig_k = eps.copy(), sig.copy()
                    CT, eps_rhs = m.tangent_phase(sig_k, self.T, self.T0, dt_cur, th)
                    self.u = self._solve(CT, eps_rhs, t)
                    eps = strain(self.coords, self.cells, self.u)
                    sig = m.post_phase(eps, sig_k, self.T, dt_cur, th)
                    if th == 1.0 or not m.elems:
                        err = 0.0
                    else:
                        err = oc.newton_error(eps_k, eps)
                    ite += 1
                    if np.isnan(err):
                        break
                if not np.isnan(err) and err <= tol:
                    converged = True
                else:
                    cuts += 1
                    sig, eps = sig_bak.copy(), eps_bak.copy()
                    m.restore(snap)
                    if cuts <= self.max_dt_cuts:
                        dt_cur = dt_cur / 2
                    else:
                        sig_k = sig_bak.copy()
            if converged:
                m.commit(sig, sig_k, dt_cur, th)
            self.sig = sig.copy()
            self.history.append(dict(t=t, u=self.u.copy(), sig=sig.copy(), eps=eps.copy(), iters=ite, error=err,
                                     dt_used=dt_cur, converged=converged))
        return self.history
