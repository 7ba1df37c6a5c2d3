"""``LinearMomentum`` with the method surface of the reference's safeincave/MomentumEquation.py
(``LinearMomentumBase`` :36-701, ``LinearMomentum`` :707-1028), re-hosted on the device.

Every method keeps its reference name, argument meaning and call order, but the work is
done by the CUDA library through ``Engine``: there is no UFL form, no sparse matrix and no PETSc.
Per-cell tensors that the reference passes around BY VALUE as CPU torch ``(N,3,3)`` tensors are
``CellField`` handles to device buffers here; user hooks that need numbers call ``.to_tensor()``
(or read ``mat.elems_ne[i].eps_ne_k`` etc., which pull lazily).

User extension point 1 is kept: subclass, override ``initialize()`` / ``run_after_solve()``.
"""
from __future__ import annotations

import numpy as np
import torch as to

from . import _lib as L
from .engine import Engine
from .MaterialProps import Material, tensor_to_voigt, voigt_to_tensor
from .Solver import KSP


class CellField:
    """Handle to a per-cell symmetric-tensor field on the device, SoA ``(6, cell_stride)``."""

    def __init__(self, engine: Engine, buf: to.Tensor):
        self.engine, self.buf = engine, buf

    def clone(self):
        return CellField(self.engine, self.buf.clone())

    def voigt(self):
        """(N,6) float64 on the host."""
        return self.buf[:, :self.engine.N].t().contiguous().cpu()

    def to_tensor(self):
        """(N,3,3) float64 on the host, the reference's currency."""
        return voigt_to_tensor(self.voigt())

    def numpy(self):
        return self.to_tensor().numpy()

    @property
    def x(self):
        """``mom_eq.sig.x.array`` of the reference's DG0 3x3 Function (flat, 9 per cell; Simulators.py:1273,
        nobian/Simulation/Run.py:1500): a host copy."""
        return _DofArray(lambda: self.to_tensor().reshape(-1).numpy())

    def __array__(self, dtype=None):
        return self.numpy()


class _DofArray:
    """Minimal stand-in for ``dolfinx.fem.Function.x``: ``.array`` is a host copy."""

    def __init__(self, getter):
        self._getter = getter

    @property
    def array(self):
        return self._getter()


class _FunctionView:
    def __init__(self, name, getter):
        self.name = name
        self.x = _DofArray(getter)


class LinearMomentumBase:
    pass


class LinearMomentum(LinearMomentumBase):
    engine_cls = Engine      # the device back end (the test-suite's host emulation substitutes its own subclass)

    def __init__(self, grid, theta: float, device="cuda"):
        self.grid, self.theta = grid, float(theta)
        self.engine = self.engine_cls(grid.tetmesh.coords, grid.tetmesh.cells, device=device)
        eng = self.engine
        self.n_elems, self.n_nodes = eng.N, eng.M
        dev = eng.device
        zn = lambda: to.zeros((eng.M, 3), dtype=to.float64, device=dev)
        self.X = zn()                       # solution vector (displacement), dof = 3*node + c
        self.b_body, self.b_neumann, self.b_ext = zn(), zn(), zn()
        self.u_prescribed = zn().reshape(-1)
        self.fixed = to.zeros(3 * eng.M, dtype=to.uint8, device=dev)
        self.dinv = to.zeros((eng.M, 9), dtype=to.float64, device=dev)
        self.solver: KSP | None = None
        self.bc = None
        self.mat: Material | None = None
        self._kelvin_phi2 = -1.0            # dt(1-theta) of the last tangent phase
        self._saved_state = None
        self._elastic_tangent_live = False
        self.ksp_log = []                   # (iterations, reason, rnorm) per linear solve
        self.sig = CellField(eng, eng.sig)  # views that follow the engine's buffers
        self.eps_tot = CellField(eng, eng.eps)
        self._nodes_vol = None
        self.dist = None                    # safeincave_b200.distributed.DistContext when partitioned
        # function-space tokens and the elastic-tangent Function the reference's example hooks touch
        # (examples/mechanics/1_triaxial/main.py:13-24); see compat.py
        from .compat import Space
        self.DG0_1, self.DG0_3x3 = Space(eng.N, 1, "DG0_1"), Space(eng.N, 9, "DG0_3x3")
        self.DG0_6x6, self.V = Space(eng.N, 36, "DG0_6x6"), Space(eng.M, 3, "V")
        self._C_fun = None
        self.mg = None                      # multigrid.Multigrid, built on the first solve with PC type "mg"
        self._guess_ring, self._guess_n, self._guess_head = None, 0, 0     # last <= 4 Newton iterates of the step
        self._solves_in_step, self._last_newton_error = 0, -1.0            # for the lagged multigrid setup (Solver.KSP)
        self._last_mg_its = None            # Krylov iterations of the previous multigrid solve of this time step
        self.guess_log = []                 # (a, b, terms) per extrapolated guess when self.guess_debug
        self.guess_debug = False
        self.mg_options = {}                # nu, coarse_its, smooth_lo, coarse_lo, safety, power_its

    # ------------------------------------------------------------------ configuration
    def set_material(self, material: Material):
        self.mat = material
        material.bind(self.engine)
        self.initialize()

    def initialize(self):
        """Hook (MomentumEquation.py:785-797).  The elastic tangent is read from the material
        table on the device; nothing to copy."""
        pass

    def run_after_solve(self):
        """Hook called after each linear solve (MomentumEquation.py:510-518)."""
        pass

    @property
    def C(self):
        """The DG0 6x6 Function the reference fills in ``initialize()`` (``self.C.x.array[:] = flatten(mat.C)``,
        MomentumEquation.py:785-797).  Kept only so that user overrides of ``initialize`` keep working: the kernels
        read the elastic tangent from the material table."""
        if self._C_fun is None:
            from .compat import Function
            self._C_fun = Function(self.DG0_6x6, "C")
        return self._C_fun

    def set_T(self, T):
        self.engine.T[:self.engine.N] = to.as_tensor(T).to(self.engine.device, dtype=to.float64)
        self.Temp = T

    def set_T0(self, T0):
        self.engine.T0[:self.engine.N] = to.as_tensor(T0).to(self.engine.device, dtype=to.float64)
        self.T0 = T0

    # ---- results to the host without stalling the time loop --------------------------------------------------------
    def fields_to_host_async(self, u_host, sig_host):
        """Displacement (M,3) and stress (6,N) of the step just finished into PINNED host tensors, overlapped with the next
        time step: one device-side copy of each into staging buffers (the next step overwrites the live ones), then the
        device-to-host copies on a copy stream of their own.  ``wait_fields()`` blocks until the host tensors are complete;
        calling this again before that simply queues behind the previous transfer."""
        eng = self.engine
        if eng.device.type != "cuda":
            u_host.copy_(self.X)
            sig_host.copy_(eng.sig[:, :eng.N])
            return
        cur = to.cuda.current_stream(eng.device)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = to.cuda.Stream(eng.device)
            self._stage_u = to.empty_like(self.X)
            self._stage_sig = to.empty((6, eng.N), dtype=to.float64, device=eng.device)
            self._copy_done = None
        if self._copy_done is not None:
            cur.wait_event(self._copy_done)              # the staging buffers are free again (device-side wait)
        self._stage_u.copy_(self.X)
        self._stage_sig.copy_(eng.sig[:, :eng.N])
        staged = to.cuda.Event()
        staged.record(cur)
        with to.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(staged)
            u_host.copy_(self._stage_u, non_blocking=True)
            sig_host.copy_(self._stage_sig, non_blocking=True)
            self._copy_done = to.cuda.Event()
            self._copy_done.record(self._copy_stream)

    def wait_fields(self):
        if getattr(self, "_copy_done", None) is not None:
            self._copy_done.synchronize()

    def set_solver(self, solver):
        if not hasattr(solver, "method"):
            raise TypeError("set_solver expects safeincave_b200.Solver.KSP (the petsc4py facade)")
        self.solver = solver

    def set_boundary_conditions(self, bc):
        self.bc = bc

    def build_body_force(self, g):
        """int rho g . v dx (MomentumEquation.py:255-275): rho_e g V_e / 4 per node of each cell."""
        eng = self.engine
        rho = to.as_tensor(self.mat.density).to(eng.device, dtype=to.float64)
        w = rho * eng.vol[:eng.N] / 4.0
        gv = to.tensor([float(c) for c in g], dtype=to.float64, device=eng.device)
        self.b_body.zero_()
        for a in range(4):
            self.b_body.index_add_(0, eng.conn[a, :eng.N].long(), w[:, None] * gv[None, :])
        eng.halo_sum(self.b_body, 3)                 # several GPUs: cells are partitioned, nodes duplicated
        to.add(self.b_body, self.b_neumann, out=self.b_ext)

    # ------------------------------------------------------------------ fields
    @property
    def u(self):
        return _FunctionView("u", lambda: self.X.reshape(-1).cpu().numpy())

    def displacement(self):
        """(n_nodes, 3) displacement on the device."""
        return self.X

    def _as_field(self, value, dst):
        """Make sure the device buffer ``dst`` holds ``value`` (CellField or (N,3,3) tensor)."""
        eng = self.engine
        if isinstance(value, CellField):
            if value.buf.data_ptr() != dst.data_ptr():
                dst.copy_(value.buf)
        elif value is not None:
            v = to.as_tensor(value)
            if v.ndim == 3:
                v = tensor_to_voigt(v.double())
            dst[:, :eng.N] = v.to(eng.device, dtype=to.float64).t()

    # ------------------------------------------------------------------ linear solve
    def _linear_solve(self):
        eng, ksp = self.engine, self.solver
        if ksp is None:
            raise RuntimeError("no solver set (set_solver)")
        if not ksp.initial_guess_nonzero:
            self.X.zero_()                                   # PETSc default: zero initial guess
        x = self.X.reshape(-1)
        extrapolate = ksp.initial_guess_nonzero and ksp.guess_extrapolation
        if extrapolate:
            if self._guess_n == 0:
                self._guess_push(x)                          # the step's starting point
            elif self._guess_n >= 4:
                coef = eng.guess_extrapolate(self._guess_iterates(), x, want_coef=self.guess_debug)
                if coef is not None:
                    self.guess_log.append(coef)
        to.where(self.fixed.bool(), self.u_prescribed, x, out=x)
        rtol, atol, max_it = ksp.effective()
        if ksp.uses_multigrid(self.grid):
            res = self._linear_solve_mg(x, rtol, atol, max_it)
        else:
            eng.block_jacobi(self.dinv, self.fixed)
            res = eng.ksp_solve(ksp.method(), self.b_ext, x, self.fixed, self.dinv, rtol=rtol, atol=atol,
                                max_it=max_it, check_every=ksp.check_every, guess_nonzero=ksp.initial_guess_nonzero)
            self._record_solve(res)
        if extrapolate:
            self._guess_push(x)
        return res

    def _record_solve(self, res):
        """Keep the Krylov result where PETSc keeps it (KSP.getConvergedReason ...) and say so when a solve did not reach
        its tolerance -- PETSc would not raise either, but the reference's users see it in -ksp_converged_reason."""
        import sys
        self.solver.record(res)
        its, reason, rnorm = int(res.iterations), int(res.reason), float(res.rnorm)
        self.ksp_log.append((its, reason, rnorm))
        if reason <= 0:
            what = {-3: "the iteration limit", -9: "a non-finite residual"}.get(reason, f"reason {reason}")
            self.ksp_failures = getattr(self, "ksp_failures", 0) + 1
            if self.ksp_failures <= 5 and self.grid.mesh.comm.rank == 0:
                print(f"[KSP] linear solve stopped on {what} after {its} iterations (residual {rnorm:.3e})"
                      + ("; further messages suppressed" if self.ksp_failures == 5 else ""), file=sys.stderr)

    def last_solve_is_nan(self):
        """True when the last linear solve met a non-finite right-hand side or residual (reason -9).  The solver leaves
        the displacement untouched in that case, so the strain would not move and the Newton measure would read 0: the
        time loop must treat it as the NaN the reference's KSP would have returned (Simulators.py:437-439)."""
        return bool(self.ksp_log) and self.ksp_log[-1][1] == -9

    # Newton iterates of the current time step, newest first, for the extrapolated Krylov guess
    def _guess_push(self, x):
        if self._guess_ring is None:
            self._guess_ring = [to.empty_like(x) for _ in range(4)]
        self._guess_head = (self._guess_head + 1) % 4
        self._guess_ring[self._guess_head].copy_(x)
        self._guess_n = min(self._guess_n + 1, 4)

    def _guess_iterates(self):
        return [self._guess_ring[(self._guess_head - k) % 4] for k in range(self._guess_n)]

    def reset_guess_history(self):
        """A new time step (or a restored one) starts: its iterates do not continue the old sequence."""
        self._guess_n = 0
        self._solves_in_step, self._last_newton_error = 0, -1.0
        self._last_mg_its = None

    def _linear_solve_mg(self, x, rtol, atol, max_it):
        """CG preconditioned by a geometric-multigrid V-cycle on the grid's refinement hierarchy (csrc/mg.cu)."""
        eng, ksp = self.engine, self.solver
        if ksp.getType().lower() != "cg":
            raise NotImplementedError("PC type 'mg' is implemented for KSP type 'cg'")
        if self.mg is None:
            from .multigrid import Multigrid
            part = getattr(self.grid, "partition", None) if (self.dist is not None and self.dist.world > 1) else None
            self.mg = Multigrid(eng, self.grid.hierarchy, part=part, coarse_fixed=self._coarse_dirichlet_mask,
                                dist=self.dist if part is not None else None,
                                **{"dist_min_cells_per_rank": getattr(self.grid, "dist_min_cells_per_rank", 200_000),
                                   **self.mg_options})
        lag = ksp.mg_setup_first > 0 and self.mg.setups > 0 and not self._elastic_tangent_live \
            and self._solves_in_step >= ksp.mg_setup_first and 0.0 <= self._last_newton_error <= ksp.mg_setup_error
        if not lag:
            t = eng._tic("mg_setup")
            self.mg.setup(self.fixed, self.dinv)
            eng._toc(t)
        self._solves_in_step += 1
        t = eng._tic("mg_solve")
        # iterations launched between two looks at the device's `done` flag: the Krylov counts of a step decrease from
        # one Newton iteration to the next, so the previous count bounds the batch (a converged solve turns the rest of its
        # batch into ~170 no-op launches per iteration)
        check = ksp.mg_check_every
        if ksp.initial_guess_nonzero and self._last_mg_its is not None:
            check = max(1, min(check, self._last_mg_its))
        res = self.mg.solve(self.b_ext, x, rtol=rtol, atol=atol, max_it=min(max_it, ksp.mg_max_it),
                            check_every=check, guess_nonzero=ksp.initial_guess_nonzero,
                            time_operator=self._time_this_solve())
        self._last_mg_its = int(res.iterations)
        eng._toc(t)
        self._record_solve(res)
        return res

    def _time_this_solve(self):
        """Engine.time_operator: True times one operator launch (and one halo exchange) in EVERY multigrid solve, an
        integer n in every n-th one -- the timed iteration is launched kernel by kernel instead of as one graph."""
        every = self.engine.time_operator
        if not every:
            return False
        self._timed_solves = getattr(self, "_timed_solves", 0) + 1
        return every is True or self._timed_solves % int(every) == 1 or int(every) == 1

    def _coarse_dirichlet_mask(self, level, mesh):
        """Dirichlet mask (3 M,) of a coarse mesh of the hierarchy, from the boundary conditions' facet tags
        (MomentumBC.py:231-245 applied to that mesh)."""
        mask = np.zeros((mesh.n_nodes, 3), dtype=np.uint8)
        for bc in self.bc.dirichlet_boundaries:
            tag = self.grid.get_boundary_tag(bc.boundary_name)
            mask[np.unique(mesh.tris[mesh.tri_tags == tag]), int(bc.component)] = 1
        return mask.reshape(-1)

    def solve_elastic_response(self):
        """MomentumEquation.py:892-923."""
        self.engine.elastic_tangent()
        self._elastic_tangent_live = True
        self.reset_guess_history()
        self._linear_solve()
        self.reset_guess_history()

    def solve(self, stress_k, t, dt):
        """MomentumEquation.py:978-1028: tangent + eps_rhs, assemble, solve, run_after_solve."""
        eng = self.engine
        self._as_field(stress_k, eng.sig_k)
        eng.tangent(dt, self.theta)                          # compute_CT + compute_eps_rhs
        self._elastic_tangent_live = False
        self._kelvin_phi2 = dt * (1.0 - self.theta)
        self._linear_solve()
        self.run_after_solve()

    def compute_CT(self, stress_k, dt):
        """MomentumEquation.py:799-820 (fused with compute_eps_rhs on the device)."""
        self._as_field(stress_k, self.engine.sig_k)
        self.engine.tangent(dt, self.theta)
        self._elastic_tangent_live = False
        self._kelvin_phi2 = dt * (1.0 - self.theta)

    def compute_eps_rhs(self, dt, stress_k):
        """MomentumEquation.py:868-890: produced by compute_CT's kernel; kept for API compatibility."""
        return CellField(self.engine, self.engine.eps_rhs)

    # ------------------------------------------------------------------ post-solve phase
    def compute_total_strain(self):
        """MomentumEquation.py:326-341."""
        self.engine.post(self.X, 0.0, self.theta, self._kelvin_phi2, L.POST_STRAIN)
        return CellField(self.engine, self.engine.eps)

    def compute_elastic_stress(self, eps_e):
        """MomentumEquation.py:822-842: sigma = C : eps."""
        eng = self.engine
        self._as_field(eps_e, eng.eps)
        if not self._elastic_tangent_live:
            eng.elastic_tangent()
            self._elastic_tangent_live = True
        eng.post(None, 0.0, self.theta, self._kelvin_phi2, L.POST_STRESS)
        return CellField(eng, eng.sig)

    def compute_stress(self, eps_tot, *_):
        """MomentumEquation.py:844-866: sigma = C_T : (eps - eps_rhs)."""
        eng = self.engine
        self._as_field(eps_tot, eng.eps)
        eng.post(None, 0.0, self.theta, self._kelvin_phi2, L.POST_STRESS)
        return CellField(eng, eng.sig)

    def increment_internal_variables(self, stress, stress_k, dt):
        """MomentumEquation.py:428-443."""
        eng = self.engine
        self._as_field(stress, eng.sig)
        self._as_field(stress_k, eng.sig_k)
        eng.post(None, dt, self.theta, self._kelvin_phi2, L.POST_INCREMENT)

    def compute_eps_ne_rate(self, stress, dt):
        """MomentumEquation.py:379-395 (phi1 = dt*theta; Simulators.py:364 passes t here, T7)."""
        eng = self.engine
        self._as_field(stress, eng.sig)
        eng.post(None, dt, self.theta, self._kelvin_phi2, L.POST_RATES)

    def newton_post(self, dt, with_error=True):
        """Fused post-solve phase of one Newton iteration (Simulators.py:416-436): strain, stress,
        ISV increment, rates and the two sums of the convergence measure in ONE kernel.
        Returns the error ||eps_k - eps|| / ||eps|| (or 0.0 if not requested)."""
        eng = self.engine
        flags = L.POST_STRAIN | L.POST_STRESS | L.POST_INCREMENT | L.POST_RATES
        if with_error:
            flags |= L.POST_ERROR
        eng.post(self.X, dt, self.theta, self._kelvin_phi2, flags)
        if self.last_solve_is_nan():
            self._last_newton_error = -1.0
            return float("nan")
        if not with_error:
            return 0.0
        if self.dist is not None and self.dist.world > 1:    # cells are partitioned: plain sums over ranks
            self.dist.all_reduce_sum(eng.err_out)
        num, den = eng.err_out.tolist()                      # device -> host read of the step result
        if den == 0.0:
            err = float("nan") if num != 0.0 else 0.0
        else:
            err = float(np.sqrt(num) / np.sqrt(den))
        self._last_newton_error = err if err == err else -1.0
        return err

    def begin_iteration(self):
        """eps_tot_k <- eps_tot, stress_k <- stress (Simulators.py:407-410) by swapping buffers."""
        eng = self.engine
        eng.sig, eng.sig_k = eng.sig_k, eng.sig
        eng.eps, eng.eps_prev = eng.eps_prev, eng.eps
        eng._prob = None
        self.sig.buf, self.eps_tot.buf = eng.sig, eng.eps
        # after the swap sig_k/eps_prev hold the current values; sig/eps are scratch until newton_post

    def update_eps_ne_rate_old(self):
        """MomentumEquation.py:397-406."""
        self.engine.commit_rates()

    def update_internal_variables(self):
        """MomentumEquation.py:445-454 -- done by commit() together with the two updates below."""
        self._pending_commit = True

    def update_eps_ne_old(self, stress, stress_k, dt):
        """MomentumEquation.py:408-426.  Runs the fused commit kernel (update_internal_variables,
        update_eps_ne_rate_old and update_eps_ne_old, in the order of Simulators.py:509-517)."""
        eng = self.engine
        self._as_field(stress, eng.sig)
        self._as_field(stress_k, eng.sig_k)
        eng.commit(dt, self.theta)
        self.reset_guess_history()

    def commit(self, dt):
        self.engine.commit(dt, self.theta)
        self.reset_guess_history()

    # ------------------------------------------------------------------ dt-retry snapshot
    def save_internal_state(self):
        """MomentumEquation.py:456-477."""
        self._saved_state = [e.snapshot() for e in self.engine.elems]

    def restore_internal_state(self):
        """MomentumEquation.py:479-494."""
        for e, s in zip(self.engine.elems, self._saved_state):
            e.restore(s)
        self.reset_guess_history()

    # ------------------------------------------------------------------ p / q output fields
    def _pq_fields(self):
        """MomentumEquation.py:287-324, 944-976 with the smoother of Grid.py:198-242, on the device (csrc/fields.cu):
        all four fields come out of one call (four small kernels; output only, outside the timed path)."""
        import ctypes
        eng = self.engine
        dev = eng.device
        if self._nodes_vol is None:
            self._nodes_vol = to.zeros(max(eng.M, 1), dtype=to.float64, device=dev)
            L.check(eng.lib.sic_node_volumes(eng._pp(), ctypes.c_void_p(self._nodes_vol.data_ptr()), eng._ph(), eng._stream()),
                    "sic_node_volumes")
            self._pq_buf = [to.zeros(max(n, 1), dtype=to.float64, device=dev) for n in (eng.M, eng.M, eng.N, eng.N)]
        pn, qn, pe, qe = self._pq_buf
        ptr = lambda t: ctypes.c_void_p(t.data_ptr())
        L.check(eng.lib.sic_pq_fields(eng._pp(), ptr(self._nodes_vol), ptr(pn), ptr(qn), ptr(pe), ptr(qe), eng._ph(),
                                      eng._stream()), "sic_pq_fields")
        eng.launches += 4
        self.p_nodes, self.q_nodes = pn[:eng.M], qn[:eng.M]
        self.p_elems, self.q_elems = pe[:eng.N], qe[:eng.N]

    def compute_p_nodes(self):
        self._pq_fields()

    def compute_q_nodes(self):
        self._pq_fields()

    def compute_p_elems(self):
        self._pq_fields()

    def compute_q_elems(self):
        self._pq_fields()
