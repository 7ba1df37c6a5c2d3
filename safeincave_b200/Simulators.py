"""``Simulator_M`` with the constructor / ``run()`` surface of the reference's
safeincave/Simulators.py:273-541, re-hosted so that the loop body stays on the device.

Loop structure, tolerances (1e-8, 40 iterations), the dt-halving retry (<= 3 cuts, T13: a retry keeps
the step's target time and BC values), the NaN exit and the "commit only if converged" rule are the
reference's.  What differs is where the data lives: ``eps_tot_k <- eps_tot`` / ``stress_k <- stress``
are buffer swaps, the post-solve phase is one fused kernel, and the only per-iteration host traffic
is the two doubles of the convergence measure.
"""
from __future__ import annotations

import sys
import time

import numpy as np


class Simulator:
    def run(self):
        raise NotImplementedError


class Simulator_M(Simulator):
    def __init__(self, eq_mom, t_control, outputs, compute_elastic_response: bool = True, verbose: bool = True):
        self.eq_mom = eq_mom
        self.t_control = t_control
        self.outputs = outputs if outputs is not None else []
        self.compute_elastic_response = compute_elastic_response
        self.verbose = verbose
        self.screen = None       # ScreenOutput.ScreenPrinter, created by run() when verbose (Simulators.py:307-308)
        self.tol, self.maxiter, self.max_dt_cuts = 1e-8, 40, 3
        self.history = []        # one dict per time step: iterations, error, converged, ksp iterations, seconds
        # optional callable(eq_mom, stress) run right after the initial stress is known and before the
        # initial rates (where the reference's scripts call desai.compute_initial_hardening between stages)
        self.after_initial_stress = None

    # -- hook for benchmarking / tests
    def on_step_end(self, record):
        pass

    def _print(self, *a):
        if self.verbose and self.eq_mom.grid.mesh.comm.rank == 0:
            print(*a)
            sys.stdout.flush()

    def initialize(self):
        """Everything of run() before the time loop (Simulators.py:338-372)."""
        eq, tc = self.eq_mom, self.t_control
        for output in self.outputs:
            output.initialize()
        eq.bc.update_dirichlet(tc.t)
        eq.bc.update_neumann(tc.t)
        if self.compute_elastic_response:
            eq.solve_elastic_response()
            eps = eq.compute_total_strain()
            stress = eq.compute_elastic_stress(eps)
        else:
            eq.compute_total_strain()
            stress = eq.sig
        if self.after_initial_stress is not None:
            self.after_initial_stress(eq, stress)
        eq.compute_eps_ne_rate(stress, tc.t)         # NB passes t, not dt (T7)
        eq.update_eps_ne_rate_old()
        self._save(0)

    def _save(self, t):
        if not self.outputs:
            return
        eq = self.eq_mom
        eq.compute_p_elems()
        eq.compute_q_elems()
        eq.compute_p_nodes()
        eq.compute_q_nodes()
        for output in self.outputs:
            output.save_fields(t)

    def step(self):
        """One pass of the time loop body (Simulators.py:378-536)."""
        eq, tc = self.eq_mom, self.t_control
        eng = eq.engine
        t0 = time.perf_counter()
        tc.advance_time()
        t, dt = tc.t, tc.dt
        sig_bak, eps_bak = eng.sig.clone(), eng.eps.clone()
        # the reference starts every KSP solve from zero, so a failed attempt cannot leak into the retry; with a warm
        # start (KSP.setInitialGuessNonzero / setGuessExtrapolation) the displacement is part of what a retry restores
        X_bak = eq.X.clone() if (eq.solver is not None and eq.solver.initial_guess_nonzero) else None
        eq.save_internal_state()
        n_ne = len(eq.mat.elems_ne)
        dt_current, dt_cut, converged = dt, 0, False
        ite, error, ksp_its = 0, 0.0, 0
        while not converged and dt_cut <= self.max_dt_cuts:
            eq.bc.update_dirichlet(t)
            eq.bc.update_neumann(t)
            error, ite = 2 * self.tol, 0
            while error > self.tol and ite < self.maxiter:
                eq.begin_iteration()                       # eps_tot_k <- eps_tot ; stress_k <- stress
                eq.solve(None, t, dt_current)
                ksp_its += eq.ksp_log[-1][0]
                want_err = not (eq.theta == 1.0 or n_ne == 0)
                error = eq.newton_post(dt_current, with_error=want_err)
                ite += 1
                if np.isnan(error):
                    break
            if not np.isnan(error) and error <= self.tol:
                converged = True
            else:
                dt_cut += 1
                eng.sig.copy_(sig_bak)
                eng.eps.copy_(eps_bak)
                eq.restore_internal_state()
                if X_bak is not None:
                    eq.X.copy_(X_bak)
                if dt_cut <= self.max_dt_cuts:
                    print(f"[SOLVER] Step {tc.step_counter}: {'NaN' if np.isnan(error) else 'no convergence'} "
                          f"after {ite} iters — halving dt ({dt_current / tc.time_conversion:.4f} -> "
                          f"{dt_current / 2 / tc.time_conversion:.4f} hours), retry {dt_cut}/{self.max_dt_cuts}",
                          file=sys.stderr)
                    dt_current = dt_current / 2
                else:
                    eng.sig_k.copy_(sig_bak)
                    dump = self._dump_nan_diagnostic(t, dt_current)
                    print(f"[SOLVER] All {self.max_dt_cuts} retries failed at step {tc.step_counter} "
                          f"(t={t / tc.time_conversion:.1f}h); state restored, step not committed. "
                          f"Diagnostic saved to {dump}", file=sys.stderr)
        if converged:
            eq.commit(dt_current)      # update_internal_variables; update_eps_ne_rate_old; update_eps_ne_old
        rec = dict(step=tc.step_counter, t=t, dt=dt, dt_used=dt_current, iterations=ite, error=float(error),
                   converged=converged, ksp_iterations=ksp_its, seconds=time.perf_counter() - t0)
        self.history.append(rec)
        self.on_step_end(rec)
        return rec

    nan_diagnostic = True     # write nan_diagnostic.pt when every dt-retry of a step failed (Simulators.py:463-503)

    def _dump_nan_diagnostic(self, t, dt):
        """Simulators.py:476-498: the restored state of the failed step as a ``torch.save`` dictionary with the
        reference's keys, pulled from the device (``G``/``B`` per element are not stored on the device -- the commit
        kernel recomputes them -- so the dump holds the material's total tangent C_T instead of ``G_total``)."""
        if not self.nan_diagnostic:
            return None
        import os
        import torch as to
        eq, eng = self.eq_mom, self.eq_mom.engine
        rank = eq.grid.mesh.comm.rank
        path = os.path.join(os.getcwd(), "nan_diagnostic.pt" if rank == 0 else f"nan_diagnostic_rank{rank}.pt")
        diag = {"step": self.t_control.step_counter, "t": t, "dt": dt, "stress": eq.sig.to_tensor(),
                "stress_backup": eq.sig.to_tensor(), "eps_tot": eq.eps_tot.to_tensor()}
        for idx, e in enumerate(eq.mat.elems_ne):
            prefix = f"elem_{idx}_{e.name}"
            diag[f"{prefix}_eps_ne_rate"] = e.eps_ne_rate
            if hasattr(e, "alpha"):
                for k in ("alpha", "alpha_0", "h", "r", "Fvp", "qsi"):
                    diag[f"{prefix}_{k}"] = getattr(e, k)
        diag["C_inv"] = eq.mat.C_inv
        diag["CT"] = to.as_tensor(eng.get_CT())
        to.save(diag, path)
        return path

    def run(self):
        tc = self.t_control
        if self.verbose:        # report header, time-step table and log.txt as the reference prints them
            from .ScreenOutput import ScreenPrinter
            ScreenPrinter.reset_instance()
            self.screen = ScreenPrinter(self.eq_mom.grid, self.eq_mom.solver, self.eq_mom.mat, self.outputs, tc.time_unit)
        self.initialize()
        while tc.keep_looping():
            rec = self.step()
            self._save(rec["t"])
            if self.screen is not None:     # Simulators.py:527-536
                current_time = "%.3f" % (rec["t"] / tc.time_conversion)
                self.screen.print_row([tc.step_counter, tc.dt / tc.time_conversion,
                                       f"{current_time} / {tc.t_final / tc.time_conversion}", rec["iterations"], rec["error"]])
        if self.screen is not None:
            self.screen.close()
        for output in self.outputs:
            if hasattr(output, "save_mesh"):
                output.save_mesh()
        return self.history


class Simulator_TM(Simulator):
    """Thermo-mechanical simulator with the constructor / ``run()`` surface of the reference's Simulator_TM
    (safeincave/Simulators.py:38-270): every step solves the heat equation, hands the cell temperatures to the
    momentum equation (thermal strain + thermally activated creep) and iterates the momentum step.  As in the
    reference: tolerance 1e-6, at most 20 iterations, NO dt-retry and an unconditional commit.  The
    temperature never leaves the device."""

    def __init__(self, eq_mom, eq_heat, t_control, outputs, compute_elastic_response: bool = True, verbose: bool = True):
        self.eq_mom, self.eq_heat = eq_mom, eq_heat
        self.t_control = t_control
        self.outputs = outputs if outputs is not None else []
        self.compute_elastic_response = compute_elastic_response
        self.verbose = verbose
        self.tol, self.maxiter = 1e-6, 20
        self.history = []

    def _save(self, t):
        if not self.outputs:
            return
        eq = self.eq_mom
        eq.compute_p_elems()
        eq.compute_q_elems()
        eq.compute_p_nodes()
        eq.compute_q_nodes()
        for output in self.outputs:
            output.save_fields(t)

    def initialize(self):
        """Simulators.py:131-176."""
        eq, heat, tc = self.eq_mom, self.eq_heat, self.t_control
        for output in self.outputs:
            output.initialize()
        eq.set_T0(heat.get_T_elems())
        eq.bc.update_dirichlet(tc.t)
        eq.bc.update_neumann(tc.t)
        if self.compute_elastic_response:
            eq.solve_elastic_response()
            eps = eq.compute_total_strain()
            stress = eq.compute_elastic_stress(eps)
        else:
            eq.compute_total_strain()
            stress = eq.sig
        T_elems = heat.get_T_elems()
        eq.set_T(T_elems)
        eq.set_T0(T_elems)
        eq.compute_eps_ne_rate(stress, tc.t)         # passes t, not dt (T7)
        eq.update_eps_ne_rate_old()
        self._save(0)

    def step(self):
        """One pass of the time loop body (Simulators.py:179-246)."""
        eq, heat, tc = self.eq_mom, self.eq_heat, self.t_control
        t0 = time.perf_counter()
        tc.advance_time()
        t, dt = tc.t, tc.dt
        eq.bc.update_dirichlet(t)
        eq.bc.update_neumann(t)
        heat.solve(t, dt)                            # updates its own BCs (HeatEquation.py:306)
        eq.set_T(heat.get_T_elems())
        n_ne = len(eq.mat.elems_ne)
        error, ite, ksp_its = 2 * self.tol, 0, 0
        while error > self.tol and ite < self.maxiter:
            eq.begin_iteration()
            eq.solve(None, t, dt)
            ksp_its += eq.ksp_log[-1][0]
            want_err = not (eq.theta == 1.0 or n_ne == 0)
            error = eq.newton_post(dt, with_error=want_err)
            ite += 1
        eq.commit(dt)
        rec = dict(step=tc.step_counter, t=t, dt=dt, iterations=ite, error=float(error), ksp_iterations=ksp_its,
                   heat_iterations=heat.ksp_log[-1][0], seconds=time.perf_counter() - t0)
        self.history.append(rec)
        return rec

    def run(self):
        tc = self.t_control
        screen = None
        if self.verbose:        # Simulators.py:89-90: the momentum equation's grid, solver and material are reported
            from .ScreenOutput import ScreenPrinter
            ScreenPrinter.reset_instance()
            screen = self.screen = ScreenPrinter(self.eq_mom.grid, self.eq_mom.solver, self.eq_mom.mat, self.outputs,
                                                 tc.time_unit)
        self.initialize()
        while tc.keep_looping():
            rec = self.step()
            self._save(rec["t"])
            if screen is not None:          # Simulators.py:257-265
                current_time = "%.3f" % (rec["t"] / tc.time_conversion)
                screen.print_row([tc.step_counter, tc.dt / tc.time_conversion,
                                  f"{current_time} / {tc.t_final / tc.time_conversion}", rec["iterations"], rec["error"]])
        if screen is not None:
            screen.close()
        for output in self.outputs:
            if hasattr(output, "save_mesh"):
                output.save_mesh()
        return self.history


class Simulator_T(Simulator):
    """Heat-diffusion-only simulator with the surface of the reference's Simulator_T (safeincave/Simulators.py:543-640):
    every step updates the heat boundary conditions, solves the backward-Euler heat equation (csrc/heat.cu through
    HeatDiffusion.solve) and saves the fields.  The step table reports 0 iterations / 0 error, as the reference's does."""

    def __init__(self, eq_heat, t_control, outputs, compute_elastic_response: bool = True, verbose: bool = True):
        self.eq_heat = eq_heat
        self.t_control = t_control
        self.outputs = outputs if outputs is not None else []
        self.verbose = verbose
        self.screen = None
        self.history = []

    def step(self):
        heat, tc = self.eq_heat, self.t_control
        t0 = time.perf_counter()
        tc.advance_time()
        t, dt = tc.t, tc.dt
        heat.bc.update_dirichlet(t)
        heat.bc.update_neumann(t)
        heat.solve(t, dt)
        rec = dict(step=tc.step_counter, t=t, dt=dt, iterations=0, error=0.0, heat_iterations=heat.ksp_log[-1][0],
                   seconds=time.perf_counter() - t0)
        self.history.append(rec)
        return rec

    def run(self):
        tc = self.t_control
        if self.verbose:
            from .ScreenOutput import ScreenPrinter
            ScreenPrinter.reset_instance()
            self.screen = ScreenPrinter(self.eq_heat.grid, self.eq_heat.solver, getattr(self.eq_heat, "mat", None), self.outputs,
                                        tc.time_unit)
        for output in self.outputs:
            output.initialize()
        for output in self.outputs:
            output.save_fields(0)
        while tc.keep_looping():
            rec = self.step()
            for output in self.outputs:
                output.save_fields(rec["t"])
            if self.screen is not None:
                current_time = "%.3f" % (rec["t"] / tc.time_conversion)
                self.screen.print_row([tc.step_counter, tc.dt / tc.time_conversion,
                                       f"{current_time} / {tc.t_final / tc.time_conversion}", 0, 0])
        if self.screen is not None:
            self.screen.close()
        for output in self.outputs:
            if hasattr(output, "save_mesh"):
                output.save_mesh()
        return self.history
