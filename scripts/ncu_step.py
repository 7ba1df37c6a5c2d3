"""One time step of the bench workload with the profiler range limited to the kernels of interest, for
`ncu --profile-from-start off --set full -k regex:...` captures (profiles/README.md holds the command lines).

    python scripts/ncu_step.py [--levels 3] [--elements dislocation] [--theta 0] [--what step|constitutive]

what=step          profile one Newton iteration of a warm time step: k_tangent, multigrid setup is OUTSIDE the range,
                   one Krylov iteration (k_mg_ebe_dot + one V-cycle: k_mg_ebe on every level, smoother, transfers), k_post, k_commit
what=constitutive  profile k_tangent / k_post / k_commit only (any element set, e.g. BASELINE config 3's)
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--levels", type=int, default=3)
    ap.add_argument("--elements", default="dislocation")
    ap.add_argument("--theta", type=float, default=0.0)
    ap.add_argument("--what", default="step", choices=["step", "constitutive"])
    ap.add_argument("--compressed", type=int, default=1)
    ap.add_argument("--staged", action="store_true", help="BASELINE config 3 through the two-stage workflow "
                    "(cases.staged_cavern_cases; block-Jacobi BiCGStab); implies --what constitutive")
    a = ap.parse_args()
    import torch
    import safeincave_b200 as sf
    from safeincave_b200 import cases
    from safeincave_b200.mesh import TetMesh
    from safeincave_b200.multigrid import refine_hierarchy
    dev = torch.device("cuda:0")
    h = refine_hierarchy(TetMesh.load_npz(os.path.join(ROOT, "tests", "golden", "mesh_cavern_regular.npz")), a.levels, device=dev)
    grid = sf.GridHandlerGMSH.from_hierarchy(h)
    if a.staged:
        a.what, a.elements = "constitutive", "kelvin,dislocation,pressure_solution,desai_cavern (staged)"
        case_eq, case = cases.staged_cavern_cases(grid, n_eq=2, n_op=3, rtol=1e-10)
        eq, sim = cases.build(case_eq, grid, device=dev)
        eq.solver.setInitialGuessNonzero(True)
        sim.verbose = False
        sim.run()
        sim = cases.add_operation_stage(case, eq, grid)
        print("equilibrium stage done; Desai cells with clamped alpha_0:", eq.mat.elems_ne[-1].n_disabled, flush=True)
    else:
        case = cases.cavern_case(grid, elements=tuple(a.elements.split(",")), theta=a.theta, n_steps=4, ksp_type="cg", rtol=1e-10)
        case["desai_initial_hardening"] = False
        eq, sim = cases.build(case, grid, device=dev)
        eq.solver.getPC().setType("mg")
        eq.mg_options = dict(eq.mg_options, compressed=bool(a.compressed))
        eq.solver.setGuessExtrapolation(True)
    sim.verbose = False
    sim.initialize()
    rec = sim.step()                             # warm: state, multigrid hierarchy, lambda_max
    assert rec["converged"], rec
    torch.cuda.synchronize()
    eng, rt = eq.engine, torch.cuda.cudart()
    dt = case["dt"]
    sim.t_control.advance_time()
    eq.bc.update_neumann(sim.t_control.t)
    eq.begin_iteration()
    rt.cudaProfilerStart()
    eng.tangent(dt, eq.theta)
    rt.cudaProfilerStop()
    eq._kelvin_phi2 = dt * (1.0 - eq.theta)
    eq._elastic_tangent_live = False
    if a.what == "step":
        eq.mg.setup(eq.fixed, eq.dinv)
        eq.mg.use_graph = False                  # kernel by kernel: the profiler sees ordinary launches
        torch.cuda.synchronize()
        rt.cudaProfilerStart()
        res = eq.mg.solve(eq.b_ext, eq.X.reshape(-1), rtol=1e-10, max_it=1, check_every=1, guess_nonzero=True)
        torch.cuda.synchronize()
        rt.cudaProfilerStop()
        eq._linear_solve()                       # finish the solve so that the post phase sees a real displacement
    else:
        eq._linear_solve()
    torch.cuda.synchronize()
    rt.cudaProfilerStart()
    err = eq.newton_post(dt)
    eq.commit(dt)
    torch.cuda.synchronize()
    rt.cudaProfilerStop()
    print(f"NCU_STEP_OK cells={eng.N} elements={a.elements} newton_error={err:.3e}")


if __name__ == "__main__":
    main()
