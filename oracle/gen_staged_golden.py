"""Golden output of the ORACLE for the staged BASELINE config-3 run on grids/cavern_irregular_finemesh (91 896 cells):
equilibrium stage (Spring + Kelvin + DislocationCreep + PressureSolutionCreep), compute_initial_hardening on the
equilibrium stress, operation stage with ViscoplasticDesai (cases.staged_cavern_cases; reference workflow
Simulators.py:1089-1326, nobian/Simulation/Run.py:1399-1510).  The oracle's sparse LU needs minutes per stage on this
grid, so its result is committed instead of recomputed by the test:

    python oracle/gen_staged_golden.py            # writes tests/golden/staged_cfg3_cavern_irregular_finemesh.npz

Stored: displacement (full), stress / strain on every 16th cell + their max-norms and 2-norms, Desai alpha_0 / alpha
(full), Newton iteration counts of every step.  Test infrastructure only.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(name="cavern_irregular_finemesh", n_eq=2, n_op=2):
    import safeincave_b200 as sf
    from safeincave_b200 import cases
    from safeincave_b200.mesh import TetMesh
    from tests.case_oracle import oracle_staged_run
    tm = TetMesh.load_npz(os.path.join(ROOT, "tests", "golden", f"mesh_{name}.npz"))
    grid = sf.GridHandlerGMSH.from_mesh(tm)
    case_eq, case_op = cases.staged_cavern_cases(grid, n_eq=n_eq, n_op=n_op)
    t0 = time.time()
    tm = grid.tetmesh            # the Morton-ordered mesh the grid (and the GPU run) uses -- NOT the file order
    osim_eq, h_eq, osim, h_op = oracle_staged_run(case_eq, case_op, tm)
    for h in h_eq[1:] + h_op[1:]:
        print(h["iters"], h["error"], h["converged"], h["dt_used"])
        assert h["converged"] and np.isfinite(h["u"]).all()
    assert all(h["dt_used"] == case_eq["dt"] for h in h_eq[1:]) and all(h["dt_used"] == case_op["dt"] for h in h_op[1:])
    last, desai = h_op[-1], osim.mat.elems[-1]
    sel = np.arange(0, tm.n_cells, 16)
    out = os.path.join(ROOT, "tests", "golden", f"staged_cfg3_{name}.npz")
    np.savez_compressed(
        out, n_eq=n_eq, n_op=n_op, n_cells=tm.n_cells,
        iters_eq=np.array([h["iters"] for h in h_eq[1:]]), iters_op=np.array([h["iters"] for h in h_op[1:]]),
        u_eq=h_eq[-1]["u"], u=last["u"], cell_sel=sel, sig_sel=last["sig"][sel], eps_sel=last["eps"][sel],
        sig_absmax=np.abs(last["sig"]).max(), eps_absmax=np.abs(last["eps"]).max(), sig_norm=np.linalg.norm(last["sig"]),
        alpha_0=desai.alpha_0, alpha=desai.alpha, n_yielding=int((desai.Fvp > 0).sum()), n_disabled=desai.n_disabled)
    print(f"wrote {out} ({os.path.getsize(out) / 1e6:.1f} MB) in {time.time() - t0:.0f} s; yielding cells "
          f"{int((desai.Fvp > 0).sum())}, clamped alpha_0 {desai.n_disabled}")


def main_cfg2(levels=1, n_steps=2):
    """BASELINE config 2 (cavern_regular, cyclic gas pressure, theta = 0, Spring + DislocationCreep, dt = 2 h) on the
    grid red-refined `levels` times -- the bench workload at the size the oracle's sparse LU still finishes in minutes:

        python oracle/gen_staged_golden.py cfg2          # writes tests/golden/cfg2_cavern_regular_L1.npz

    The mesh is the finest level of multigrid.refine_hierarchy (what the multigrid tests and bench.py build)."""
    import safeincave_b200 as sf
    from safeincave_b200 import cases
    from safeincave_b200.mesh import TetMesh
    from safeincave_b200.multigrid import refine_hierarchy
    from tests.case_oracle import oracle_simulator
    h = refine_hierarchy(TetMesh.load_npz(os.path.join(ROOT, "tests", "golden", "mesh_cavern_regular.npz")), levels)
    tm = h.finest
    grid = sf.GridHandlerGMSH.from_hierarchy(h)
    case = cases.cavern_case(grid, n_steps=n_steps, ksp_type="cg", rtol=1e-12)
    t0 = time.time()
    hist = oracle_simulator(case, tm).run(0.0, [case["dt"]] * n_steps)
    for r in hist[1:]:
        print(r["iters"], r["error"], r["converged"], r["dt_used"])
        assert r["converged"] and r["dt_used"] == case["dt"] and np.isfinite(r["u"]).all()
    last = hist[-1]
    sel = np.arange(0, tm.n_cells, 16)
    out = os.path.join(ROOT, "tests", "golden", f"cfg2_cavern_regular_L{levels}.npz")
    np.savez_compressed(out, levels=levels, n_steps=n_steps, n_cells=tm.n_cells, iters=np.array([r["iters"] for r in hist[1:]]),
                        u=last["u"], cell_sel=sel, sig_sel=last["sig"][sel], eps_sel=last["eps"][sel],
                        sig_absmax=np.abs(last["sig"]).max(), eps_absmax=np.abs(last["eps"]).max(),
                        sig_norm=np.linalg.norm(last["sig"]), coords_checksum=float(np.abs(tm.coords).sum()))
    print(f"wrote {out} ({os.path.getsize(out) / 1e6:.1f} MB) in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    if sys.argv[1:2] == ["cfg2"]:
        main_cfg2()
    else:
        main(*sys.argv[1:2])
