"""CPU oracle of the geometric-multigrid preconditioned CG (csrc/mg.cu).

TEST INFRASTRUCTURE ONLY (tests/ and scripts/ import it; the product never does).

The reference has no counterpart to pin this against: its preconditioner is whatever PETSc PC the user
script picks (MomentumEquation.py:1023-1025; `asm` in the examples, `gamg` in
nobian/run_interlayer.py:2114-2116), i.e. third-party code outside /root/reference -- PARITY UNPINNED in
that sense.  What IS pinned: the preconditioner only changes the path of the Krylov iteration, not its
fixed point, so every MG solve is checked against the sparse direct solve of oracle/fem.py; and this
module restates the cycle with ASSEMBLED scipy matrices (explicit P, K_l, D_l^-1), a formulation
independent of the matrix-free kernels, so that one V-cycle can be compared vector for vector.

Algorithm (same parameters as sic_mg_opts_t): V(nu,nu), smoother = Chebyshev polynomial in D^-1 K for
eigenvalues in [smooth_lo, 1] * lambda_max with D the nodal 3x3 block diagonal, coarsest level =
`coarse_its` Chebyshev steps on [coarse_lo, 1] * lambda_max, Galerkin coarse operators = operators of the
mean of the eight children's C_T (identity checked in tests/test_oracle_mg.py), Dirichlet dofs kept zero.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import fem


def block_diag_inverse(K, n_nodes):
    blk = np.zeros((n_nodes, 3, 3))
    Kc = K.tocoo()
    sel = (Kc.row // 3) == (Kc.col // 3)
    np.add.at(blk, (Kc.row[sel] // 3, Kc.row[sel] % 3, Kc.col[sel] % 3), Kc.data[sel])
    bi = np.linalg.inv(blk)
    n = np.arange(n_nodes)
    r = np.broadcast_to(3 * n[:, None, None] + np.arange(3)[None, :, None], (n_nodes, 3, 3))
    c = np.broadcast_to(3 * n[:, None, None] + np.arange(3)[None, None, :], (n_nodes, 3, 3))
    return sp.csr_matrix((bi.ravel(), (r.ravel(), c.ravel())), shape=K.shape)


def eliminated(K, fixed):
    """assemble_matrix(bcs): rows/columns of fixed dofs zeroed, unit diagonal (MomentumEquation.py:1010)."""
    free = (~fixed).astype(float)
    D = sp.diags(free)
    return (D @ K @ D + sp.diags(fixed.astype(float))).tocsr()


class OracleMG:
    def __init__(self, meshes, transfers, CT_fine, fixed_fine, nu=2, coarse_its=30, smooth_lo=0.1, coarse_lo=0.01):
        """meshes / transfers: safeincave_b200.multigrid.Hierarchy fields; CT_fine (N,6,6); fixed_fine bool (3M,)."""
        from safeincave_b200.multigrid import prolongation_matrix
        n = len(meshes)
        self.nu, self.coarse_its, self.smooth_lo, self.coarse_lo = nu, coarse_its, smooth_lo, coarse_lo
        self.CT, self.fixed = [None] * n, [None] * n
        self.CT[-1], self.fixed[-1] = np.asarray(CT_fine), np.asarray(fixed_fine, dtype=bool)
        for l in range(n - 1, 0, -1):
            t = transfers[l]
            self.CT[l - 1] = self.CT[l][t.children.astype(np.int64)].mean(axis=0)      # (8,Nc,6,6) -> (Nc,6,6)
            self.fixed[l - 1] = self.fixed[l].reshape(-1, 3)[t.inject.astype(np.int64)].reshape(-1)
        self.K, self.Dinv, self.P = [], [], [None]
        for l, m in enumerate(meshes):
            K = eliminated(fem.assemble_K(m.coords, m.cells, self.CT[l]), self.fixed[l])
            self.K.append(K)
            self.Dinv.append(block_diag_inverse(K, m.n_nodes))
        for l in range(1, n):
            Pn = prolongation_matrix(transfers[l], meshes[l - 1].n_nodes)
            P = sp.kron(Pn, sp.identity(3), format="csr")
            P = sp.diags((~self.fixed[l]).astype(float)) @ P @ sp.diags((~self.fixed[l - 1]).astype(float))
            self.P.append(P.tocsr())
        self.lam = [None] * n

    def power_lambda(self, l, its=30, seed=1):
        rng = np.random.default_rng(seed)
        v = rng.standard_normal(self.K[l].shape[0]) * (~self.fixed[l])
        v /= np.linalg.norm(v)
        lam = 0.0
        for _ in range(its):
            w = (self.Dinv[l] @ (self.K[l] @ v)) * (~self.fixed[l])
            lam = np.linalg.norm(w)
            v = w / lam
        return lam

    def chebyshev(self, l, b, x, its, lo):
        lmax = self.lam[l]
        lmin = lo * lmax
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma = theta / delta
        rho = 1.0 / sigma
        free = ~self.fixed[l]
        if x is None:
            r, x = b * free, np.zeros_like(b)
        else:
            r = (b - self.K[l] @ x) * free
        d = (self.Dinv[l] @ r) * free / theta
        x = x + d
        for _ in range(1, its):
            r = (r - self.K[l] @ d) * free
            rho_new = 1.0 / (2 * sigma - rho)
            d = (rho_new * rho * d + (2 * rho_new / delta) * (self.Dinv[l] @ r)) * free
            x = x + d
            rho = rho_new
        return x, r, d

    def vcycle(self, l, b):
        if l == 0:
            return self.chebyshev(0, b, None, self.coarse_its, self.coarse_lo)[0]
        free = ~self.fixed[l]
        x, r, d = self.chebyshev(l, b, None, self.nu, self.smooth_lo)
        r = (r - self.K[l] @ d) * free                       # true residual of x
        xc = self.vcycle(l - 1, self.P[l].T @ r)
        x = x + self.P[l] @ xc
        return self.chebyshev(l, b, x, self.nu, self.smooth_lo)[0]

    def pcg(self, b, x0=None, rtol=1e-12, max_it=500):
        top = len(self.K) - 1
        A = self.K[top]
        x = np.zeros_like(b) if x0 is None else x0.copy()
        r = (b - A @ x) * (~self.fixed[top])
        r0 = np.linalg.norm(r)
        z = self.vcycle(top, r)
        p, rz = z.copy(), r @ z
        for it in range(1, max_it + 1):
            q = A @ p
            a = rz / (p @ q)
            x += a * p
            r -= a * q
            if np.linalg.norm(r) <= rtol * r0:
                return x, it
            z = self.vcycle(top, r)
            rz2 = r @ z
            p = z + (rz2 / rz) * p
            rz = rz2
        return x, max_it
