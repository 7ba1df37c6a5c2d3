"""Parity checks of the heat equation and of the thermo-mechanical loop, shared by the
host-emulation tests (tests/test_emu_heat.py) and the B200 tests (tests/test_gpu_heat.py)."""
import os

import numpy as np
import torch as to

from oracle import heat as oh
from tests.case_oracle import oracle_simulator

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DAY = 86400.0


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def load_grid(sf, name, levels=0):
    from safeincave_b200.mesh import TetMesh, red_refine
    tm = TetMesh.load_npz(os.path.join(GOLD, f"mesh_{name}.npz"))
    for _ in range(levels):
        tm = red_refine(tm)
    return sf.GridHandlerGMSH.from_mesh(tm)


def heat_case(grid, name):
    if name == "cavern_overburden_coarse":          # BASELINE config 4: thermomechanics/2_cavern/main.py:217-272
        from safeincave_b200 import cases
        return cases.overburden_tm_case(grid)["thermal"]
    return _heat_case(grid, name)


def _heat_case(grid, name):
    """Boundary data in the pattern of examples/thermomechanics/2_cavern/main.py:243-273 (Dirichlet top, geothermal
    flux at the bottom, Robin h = 5 W/m2/K towards the gas on the cavern wall; cp 850, k 7)."""
    names = {n.upper(): n for n in grid.get_boundary_names()}
    x = grid.mesh.geometry.x
    z_top = float(x[:, 2].max())
    if name == "cube_coarse":
        T0 = 300.0 + 20.0 * (z_top - x[:, 2])
        return dict(T0=T0, t_final=10 * DAY, rho=2200.0, cp=850.0, k=7.0,
                    dirichlet=[dict(boundary=names["TOP"], values=[300.0, 310.0], time_values=[0.0, 10 * DAY])],
                    neumann=[dict(boundary=names["BOTTOM"], values=[0.5, 0.5], time_values=[0.0, 10 * DAY])],
                    robin=[dict(boundary=names["EAST"], h=5.0, values=[290.0, 270.0], time_values=[0.0, 10 * DAY])])
    T0 = 293.0 + 0.027 * (z_top - x[:, 2])
    return dict(T0=T0, t_final=240 * DAY, rho=2200.0, cp=850.0, k=7.0,
                dirichlet=[dict(boundary="Top", values=[293.0, 293.0], time_values=[0.0, 240 * DAY])],
                neumann=[dict(boundary="Bottom", values=[0.027 * 7.0] * 2, time_values=[0.0, 240 * DAY])],
                robin=[dict(boundary="Cavern", h=5.0, values=[283.0, 283.0], time_values=[0.0, 240 * DAY])])


def build_heat(sf, grid, hc, rtol=1e-13):
    from safeincave_b200 import HeatBC as heatBC
    n = grid.n_elems
    heat = sf.HeatDiffusion(grid)
    ksp = sf.PETSc.KSP().create(grid.mesh.comm)
    ksp.setType("cg")
    ksp.getPC().setType("asm")
    ksp.setTolerances(rtol=rtol, max_it=100)
    heat.set_solver(ksp)
    mat = sf.Material(n)
    from safeincave_b200.cases import per_cell
    cellwise = lambda v: to.as_tensor(per_cell(v, grid.tetmesh, n))
    mat.set_density(cellwise(hc["rho"]))
    mat.set_specific_heat_capacity(cellwise(hc["cp"]))
    mat.set_thermal_conductivity(cellwise(hc["k"]))
    heat.set_material(mat)
    heat.set_initial_T(to.as_tensor(hc["T0"]))
    bc = heatBC.BcHandler(heat)
    for d in hc["dirichlet"]:
        bc.add_boundary_condition(heatBC.DirichletBC(d["boundary"], d["values"], d["time_values"]))
    for d in hc["neumann"]:
        bc.add_boundary_condition(heatBC.NeumannBC(d["boundary"], d["values"], d["time_values"]))
    for d in hc["robin"]:
        bc.add_boundary_condition(heatBC.RobinBC(d["boundary"], d["values"], d["h"], d["time_values"]))
    heat.set_boundary_conditions(bc)
    return heat


def oracle_heat(tm, hc):
    tag = lambda b: tm.names[2][b]
    conv = lambda lst, robin=False: [dict(tag=tag(d["boundary"]), values=d["values"], time_values=d["time_values"],
                                          **({"h": d["h"]} if robin else {})) for d in lst]
    n = tm.n_cells
    from safeincave_b200.cases import per_cell
    o = oh.OracleHeat(tm.coords, tm.cells, tm.tris, tm.tri_tags, per_cell(hc["rho"], tm), per_cell(hc["cp"], tm),
                      per_cell(hc["k"], tm), conv(hc["dirichlet"]), conv(hc["neumann"]), conv(hc["robin"], True))
    o.set_initial_T(hc["T0"])
    return o


def check_heat_steps(sf, name, levels, n_steps, dt):
    grid = load_grid(sf, name, levels)
    hc = heat_case(grid, name)
    heat = build_heat(sf, grid, hc)
    o = oracle_heat(grid.tetmesh, hc)
    for i in range(n_steps):
        t = (i + 1) * dt
        heat.solve(t, dt)
        o.step(t, dt)
        assert heat.ksp_log[-1][1] > 0
        # temperatures are ~300 K: compare the CHANGE since the start as well, so that the check has teeth
        assert relerr(heat.T.x.array, o.T) < 2e-10          # Krylov tolerance (rtol 1e-13 on ||b||, T ~ 300 K)
        assert relerr(heat.T.x.array - hc["T0"], o.T - hc["T0"]) < 1e-8
    assert np.abs(o.T - hc["T0"]).max() > 0.5, "nothing happened: the check is vacuous"
    assert relerr(heat.get_T_elems().cpu().numpy(), heat.T.x.array[grid.tetmesh.cells].mean(axis=1)) < 1e-14
    assert relerr(heat.get_T_elems().cpu().numpy(), o.cell_mean()) < 2e-10
    assert relerr(heat.T_old.x.array, o.T_old) < 2e-10
    return max(k[0] for k in heat.ksp_log)


def check_thermomechanical_steps(sf, name="cube_coarse", levels=1, n_steps=3, dt=0.5 * DAY, tol=1e-8, tol_T=2e-10):
    """Simulator_TM against OracleSimulatorTM: Spring + Thermoelastic + Kelvin + DislocationCreep; the Robin / Dirichlet
    data cool one side by tens of kelvin within the run, so thermal strain and the Arrhenius factor both move."""
    from safeincave_b200 import cases
    grid = load_grid(sf, name, levels)
    tm = grid.tetmesh
    hc = heat_case(grid, name)
    heat = build_heat(sf, grid, hc)
    case_fn = {"cube_coarse": cases.triaxial_case, "cavern_overburden_coarse": cases.overburden_tm_case}.get(name, cases.cavern_case)
    case = case_fn(grid, n_steps=n_steps)
    case["dt"] = dt
    case["t_final_run"] = n_steps * dt
    case.setdefault("thermo_alpha", 44e-6)                            # thermomechanics/2_cavern/main.py:91
    eq, sim_m = cases.build(case, grid)
    eq.mat.add_to_thermoelastic(sf.Thermoelastic(to.as_tensor(cases.per_cell(case["thermo_alpha"], tm))))
    eq.set_material(eq.mat)
    sim = sf.Simulator_TM(eq, heat, sim_m.t_control, [], compute_elastic_response=True, verbose=False)
    hist = sim.run()
    om = oracle_simulator(case, tm)
    om.mat.add_thermoelastic(cases.per_cell(case["thermo_alpha"], tm))
    osim = oh.OracleSimulatorTM(om, oracle_heat(tm, hc))
    ohist = osim.run(0.0, [dt] * n_steps)
    assert [h["iterations"] for h in hist] == [h["iters"] for h in ohist[1:]]
    last = ohist[-1]
    eng = eq.engine
    assert relerr(heat.T.x.array, last["T"]) < tol_T     # Krylov tolerance of the heat solve (rtol 1e-13 on ||b||, T ~ 300 K)
    assert relerr(eq.X.reshape(-1).cpu().numpy(), last["u"]) < tol
    assert relerr(eng.get6(eng.sig), last["sig"]) < tol
    assert relerr(eng.get1(eng.T), osim.mech.T) < tol_T
    assert relerr(eng.get1(eng.T0), osim.mech.T0) < 1e-14
    for e_gpu, e_or in zip(eng.elems, om.mat.elems):
        assert relerr(eng.get6(e_gpu.eps_old), e_or.eps_old) < tol
    # the thermal strain matters in this run: without it the displacement differs visibly
    assert np.abs(osim.mech.T - osim.mech.T0).max() > 1.0
    return hist
