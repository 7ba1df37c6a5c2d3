"""The CPU arm of bench.py: the oracle's time step with every host core put to work.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (imported by bench.py's ``cpu_baseline`` / ``--impl reference`` legs and by
tests/; never by the product).

The reference's own stack (FEniCSx/PETSc under mpirun) cannot be installed here, so the CPU baseline is
the oracle port of the same path, organised the way the reference's MPI run organises it:

    constitutive update   oracle/constitutive.py (numpy restatement of MaterialProps.py:172-309 and the element
                          classes), the cells cut into one contiguous chunk per thread -- the per-cell update is
                          embarrassingly parallel, as it is across MPI ranks in the reference;
    assembly              scipy CSR (fem.assemble_K / rhs_eps) per chunk, summed (MomentumEquation.py:1008-1020);
    solve                 preconditioned CG on the assembled matrix to the SAME relative tolerance as the GPU arm,
                          zero initial guess as PETSc's default (MomentumEquation.py:1023-1025), nodal 3x3
                          block-Jacobi preconditioner (what PETSc's ``asm`` degenerates to with one node per block;
                          ILU(0) blocks would need fewer iterations but cost a sequential triangular solve per thread),
                          SpMV split into row blocks over the threads (scipy's kernels release the GIL).

``ThreadedSimulatorM`` is OracleSimulatorM with these two substitutions; tests/test_cpu_step.py checks it against the
plain oracle (sparse LU).
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np
import scipy.sparse as sp

from . import constitutive as oc
from . import fem


def _chunks(n, k):
    b = [(n * i) // k for i in range(k + 1)]
    return [(b[i], b[i + 1]) for i in range(k) if b[i + 1] > b[i]]


class ChunkedMaterial:
    """An OracleMaterial per contiguous chunk of cells, driven by a thread pool; the interface OracleSimulatorM uses."""

    def __init__(self, make_material, n, pool, n_chunks):
        """make_material(c0, c1) -> OracleMaterial of the cells [c0, c1)."""
        self.n, self.pool = n, pool
        self.spans = _chunks(n, n_chunks)
        self.mats = list(pool.map(lambda s: make_material(*s), self.spans))
        self.C = np.concatenate([m.C for m in self.mats])
        self.elems = self.mats[0].elems            # truthiness / count only

    def _map(self, fn, *arrays):
        def run(k):
            c0, c1 = self.spans[k]
            return fn(self.mats[k], *[a[c0:c1] if isinstance(a, np.ndarray) and a.shape[:1] == (self.n,) else a for a in arrays])
        return list(self.pool.map(run, range(len(self.spans))))

    def tangent_phase(self, sig_k, T, T0, dt, theta):
        out = self._map(lambda m, s, t, t0: m.tangent_phase(s, t, t0, dt, theta), sig_k, T, T0)
        return np.concatenate([o[0] for o in out]), np.concatenate([o[1] for o in out])

    def post_phase(self, eps, sig_k, T, dt, theta):
        return np.concatenate(self._map(lambda m, e, s, t: m.post_phase(e, s, t, dt, theta), eps, sig_k, T))

    def elastic_stress(self, eps):
        return np.concatenate(self._map(lambda m, e: m.elastic_stress(e), eps))

    def eval_rates(self, sig, phi1, T):
        self._map(lambda m, s, t: m.eval_rates(s, phi1, t), sig, T)

    def commit(self, sig, sig_k, dt, theta):
        self._map(lambda m, s, sk: m.commit(s, sk, dt, theta), sig, sig_k)

    def commit_rates(self):
        self._map(lambda m: m.commit_rates())

    def snapshot(self):
        return self._map(lambda m: m.snapshot())

    def restore(self, snaps):
        for m, s in zip(self.mats, snaps):
            m.restore(s)


class ThreadedSimulatorM(fem.OracleSimulatorM):
    """OracleSimulatorM (Simulators.py:310-541) with chunked constitutive updates, chunked assembly and a threaded
    block-Jacobi PCG instead of the sparse LU."""

    def __init__(self, *a, n_threads=1, rtol=1e-10, max_it=20000, **k):
        super().__init__(*a, **k)
        self.n_threads, self.rtol, self.max_it = int(n_threads), float(rtol), int(max_it)
        self.pool = ThreadPoolExecutor(self.n_threads)
        self.krylov_iterations = []
        self._cell_spans = _chunks(self.cells.shape[0], self.n_threads)

    # ---- assembly, one chunk of cells per thread
    def _assemble(self, CT, eps_rhs):
        def part(span):
            c0, c1 = span
            return (fem.assemble_K(self.coords, self.cells[c0:c1], CT[c0:c1]),
                    fem.rhs_eps(self.coords, self.cells[c0:c1], CT[c0:c1], eps_rhs[c0:c1]))
        parts = list(self.pool.map(part, self._cell_spans))
        K, b = parts[0]
        for Kp, bp in parts[1:]:
            K = K + Kp
            b = b + bp
        return K.tocsr(), b

    def _solve(self, CT, eps_rhs, t):
        dofs, vals, b_ext = self._bc(t)
        K, b_eps = self._assemble(CT, eps_rhs)
        b = b_ext + b_eps
        n = K.shape[0]
        u = np.zeros(n)
        u[dofs] = vals
        free = np.ones(n, dtype=bool)
        free[dofs] = False
        b = b - K @ u                                   # lifting (apply_lifting + set_bc, MomentumEquation.py:1017-1020)
        b[~free] = 0.0
        # rows / columns of constrained dofs are dropped by masking vectors (assemble_matrix(bcs) keeps a unit diagonal)
        row_spans = _chunks(n, self.n_threads)
        blocks = [K[r0:r1] for r0, r1 in row_spans]

        def matvec(x):
            y = np.empty(n)

            def run(k):
                r0, r1 = row_spans[k]
                y[r0:r1] = blocks[k] @ x
            list(self.pool.map(run, range(len(blocks))))
            y[~free] = 0.0
            return y
        # nodal 3x3 block-Jacobi: invert the diagonal blocks with constrained dofs replaced by the identity
        M = n // 3
        idx = 3 * np.arange(M)
        D = np.zeros((M, 3, 3))
        Kd = K.tocsr()
        for i in range(3):
            for j in range(3):
                D[:, i, j] = np.asarray(Kd[idx + i, idx + j]).ravel()
        fm = free.reshape(M, 3)
        for i in range(3):
            off = ~fm[:, i]
            D[off, i, :] = 0.0
            D[off, :, i] = 0.0
            D[off, i, i] = 1.0
        Dinv = np.linalg.inv(D)

        def precond(r):
            z = np.einsum("nij,nj->ni", Dinv, r.reshape(M, 3)).reshape(-1)
            z[~free] = 0.0
            return z
        x = np.zeros(n)
        r = b.copy()
        bnorm = np.linalg.norm(b)
        z = precond(r)
        p = z.copy()
        rz = r @ z
        its = 0
        while np.linalg.norm(r) > self.rtol * bnorm and its < self.max_it:
            q = matvec(p)
            alpha = rz / (p @ q)
            x += alpha * p
            r -= alpha * q
            z = precond(r)
            rz_new = r @ z
            p = z + (rz_new / rz) * p
            rz = rz_new
            its += 1
        self.krylov_iterations.append(its)
        return u + x


def threaded_simulator(case, tm, n_threads, rtol=1e-10):
    """The simulator of tests/case_oracle.oracle_simulator, threaded (same case dict, same mesh)."""
    from tests.case_oracle import oracle_material
    from safeincave_b200.cases import cell_temperature
    n = tm.n_cells
    pool = ThreadPoolExecutor(n_threads)
    mat = ChunkedMaterial(lambda c0, c1: oracle_material(case, c1 - c0), n, pool, n_threads)
    T = cell_temperature(case, tm.coords, tm.cells)
    tag = lambda name: tm.names[2][name]
    dirichlet = [dict(tag=tag(d["boundary"]), component=d["component"], values=d["values"],
                      time_values=d["time_values"]) for d in case["dirichlet"]]
    neumann = [dict(tag=tag(b["boundary"]), direction=b["direction"], density=b["density"], ref_pos=b["ref_pos"],
                    gravity=b["gravity"], values=b["values"], time_values=b["time_values"]) for b in case["neumann"]]
    return ThreadedSimulatorM(tm.coords, tm.cells, tm.tris, tm.tri_tags, mat, case["theta"], T, T,
                              case["density"] * np.ones(n), case["g"], dirichlet, neumann, n_threads=n_threads, rtol=rtol)
