"""The BASELINE.json configurations as plain data + a builder for the device-side objects.

A case is a dict (no torch / CUDA objects) so that tests can build the CPU oracle from the very
same numbers:

    triaxial_case(grid)   config 1: unit-cube triaxial creep, Spring + Kelvin + DislocationCreep
                          (+ Desai), loads / BCs of examples/mechanics/1_triaxial/main.py:36-147
    cavern_case(grid)     configs 2/3/5: single-region cavern grid, cyclic gas pressure; BC pattern of
                          examples/mechanics/nobian/Simulation/Run.py:1414-1430, materials of
                          examples/mechanics/4_cavern/main.py:55-79, geothermal T (:90-97)
"""
from __future__ import annotations

import numpy as np
import torch as to

from . import MomentumBC as momBC
from .Utils import GPa, MPa, day, hour

DESAI_TRIAXIAL = dict(mu_1=5.3665857009859815e-11, N_1=3.1, a_1=1.965018496922832e-05,
                      eta=0.8275682807874163, n=3.0, beta_1=0.0048, beta=0.995, m=-0.5,
                      gamma=0.095, sigma_t=5.0, alpha_0=0.0022)         # 1_triaxial/main.py:78-89

# nobian/Simulation/Run.py:1456-1462 (scenario A, CCC Zuidwending) with the fixed shape parameters of :1486-1492
DESAI_CAVERN_A = dict(DESAI_TRIAXIAL, mu_1=6.89e-12, N_1=3.0, a_1=1.80e-5, eta=0.82, alpha_0=2.0e-3)

ELEMENT_LIBRARY = {
    "kelvin": dict(kind="kelvin", eta=105e11, E=10 * GPa, nu=0.32),                    # 1_triaxial/main.py:66-69
    "dislocation": dict(kind="dislocation", A=1.9e-20, Q=51600.0, n=3.0),              # :72-75
    "pressure_solution": dict(kind="pressure_solution", A=1.29e-19, d=0.01, Q=13184.0),  # thermomechanics/2_cavern/main.py:82-87
    "desai": dict(kind="desai", **DESAI_TRIAXIAL),
    "desai_cavern": dict(kind="desai", **DESAI_CAVERN_A),
    # nobian/Simulation/Run.py:1262-1276 (scenario A, Munson-Dawson model; mu = E0 / (2 (1 + nu0)))
    "munson_dawson": dict(kind="munson_dawson", A=18.31 * (1e-6) ** 4.99 / (365 * 24 * 3600.0), Q=6356.0 * 8.32, n=4.99,
                          K0=7.0e-7, c=9.02e-3, m=3.0, alpha_w=-13.2, beta_w=-7.738, delta=0.58, mu=20.425e9 / 2.5),
    # nobian/Simulation/run_interlayer.py:110-111, 1617-1623 (anhydrite interlayer)
    "mohr_coulomb": dict(kind="mohr_coulomb", mu_1=1e-9, N_1=1.0, cohesion=4.0, friction_angle=float(np.radians(35.0)),
                         dilation_angle=0.0, sigma_t=1.0),
    "matsuoka_nakai": dict(kind="matsuoka_nakai", mu_1=1e-9, N_1=1.0, cohesion=4.0, friction_angle=float(np.radians(35.0)),
                           dilation_angle=0.0, sigma_t=1.0),
}


def triaxial_case(grid, elements=("kelvin", "dislocation"), n_steps=None, ksp_override=None):
    t_final = 24 * hour
    names = grid.get_boundary_names()
    up = {n.upper(): n for n in names}
    case = dict(
        name="triaxial", theta=0.5, dt=0.5 * hour, t_final=t_final, time_unit="hour",
        density=2000.0, g=[0.0, 0.0, 0.0], T=293.0,
        spring=dict(E=102 * GPa, nu=0.3),
        elements=[dict(ELEMENT_LIBRARY[e]) if isinstance(e, str) else dict(e) for e in elements],
        dirichlet=[dict(boundary=up["WEST"], component=0, values=[0.0, 0.0], time_values=[0.0, t_final]),
                   dict(boundary=up["BOTTOM"], component=2, values=[0.0, 0.0], time_values=[0.0, t_final]),
                   dict(boundary=up["SOUTH"], component=1, values=[0.0, 0.0], time_values=[0.0, t_final])],
        neumann=[dict(boundary=up["EAST"], direction=2, density=0.0, ref_pos=0.0, gravity=0.0,
                      values=[4.0 * MPa, 4.0 * MPa], time_values=[0.0, t_final]),
                 dict(boundary=up["NORTH"], direction=2, density=0.0, ref_pos=0.0, gravity=0.0,
                      values=[4.0 * MPa, 4.0 * MPa], time_values=[0.0, t_final]),
                 dict(boundary=up["TOP"], direction=2, density=0.0, ref_pos=0.0, gravity=0.0,
                      values=[4.1 * MPa, 16 * MPa, 16 * MPa, 6 * MPa, 6 * MPa],
                      time_values=[0 * hour, 2 * hour, 14 * hour, 16 * hour, 24 * hour])],
        ksp=dict(type=ksp_override or "bicg", rtol=1e-12), desai_initial_hardening=False,
    )
    if n_steps is not None:
        case["t_final_run"] = n_steps * case["dt"]
    return case


def cavern_case(grid, elements=("dislocation",), theta=0.0, dt_hours=2.0, p_ref=17.5 * MPa, n_steps=None,
                ksp_type="cg", rtol=1e-12):
    x = grid.mesh.geometry.x
    z_top = float(x[:, 2].max())
    tm = grid.tetmesh
    cav_tag = grid.get_boundary_tag("Cavern")
    z_max = float(x[np.unique(tm.tris[tm.tri_tags == cav_tag]), 2].max())     # cavern roof
    salt_density, gas_density, g = 2200.0, 0.082, -9.81
    p_roof = p_ref + salt_density * abs(g) * (z_top - z_max)
    t_final = 240 * hour
    case = dict(
        name="cavern", theta=theta, dt=dt_hours * hour, t_final=t_final, time_unit="hour",
        density=salt_density, g=[0.0, 0.0, g],
        T=dict(surface=293.0, gradient=27.0 / 1000.0, z_surface=z_top),           # 4_cavern/main.py:90-97
        spring=dict(E=102 * GPa, nu=0.3),
        elements=[dict(ELEMENT_LIBRARY[e]) if isinstance(e, str) else dict(e) for e in elements],
        dirichlet=[dict(boundary="West", component=0, values=[0.0, 0.0], time_values=[0.0, t_final]),
                   dict(boundary="Bottom", component=2, values=[0.0, 0.0], time_values=[0.0, t_final]),
                   dict(boundary="South", component=1, values=[0.0, 0.0], time_values=[0.0, t_final])],
        neumann=[dict(boundary="East", direction=2, density=salt_density, ref_pos=z_top, gravity=g,
                      values=[p_ref, p_ref], time_values=[0.0, t_final]),
                 dict(boundary="North", direction=2, density=salt_density, ref_pos=z_top, gravity=g,
                      values=[p_ref, p_ref], time_values=[0.0, t_final]),
                 dict(boundary="Top", direction=2, density=0.0, ref_pos=0.0, gravity=g,
                      values=[p_ref, p_ref], time_values=[0.0, t_final]),
                 dict(boundary="Cavern", direction=2, density=gas_density, ref_pos=z_max, gravity=g,
                      values=[0.8 * p_roof, 0.2 * p_roof, 0.2 * p_roof, 0.8 * p_roof, 0.8 * p_roof],
                      time_values=[0 * day, 2 * day, 6 * day, 8 * day, 10 * day])],
        ksp=dict(type=ksp_type, rtol=rtol), desai_initial_hardening=("desai" in elements),
    )
    if n_steps is not None:
        case["t_final_run"] = n_steps * case["dt"]
    return case


def _staged(make_case, creep, desai, n_eq, n_op):
    case_eq = make_case(tuple(creep), n_eq, True)
    case_eq["desai_initial_hardening"] = False
    for bc in case_eq["neumann"] + case_eq["dirichlet"]:      # time-constant boundary values (Simulators.py:1156-1158)
        bc["values"] = [bc["values"][0]] * len(bc["values"])
    case_op = make_case(tuple(creep) + (desai,), n_op, False)
    case_op["desai_initial_hardening"] = False
    case_op["stage_elements"] = case_op["elements"][len(creep):]
    return case_eq, case_op


def staged_cavern_cases(grid, n_eq=2, n_op=2, dt_eq_hours=0.5, dt_op_hours=0.5, theta=0.5, ksp_type="bicg", rtol=1e-12,
                        creep=("kelvin", "dislocation", "pressure_solution"), desai="desai_cavern"):
    """BASELINE config 3 the way the reference runs it (Simulators.py:1089-1191 run_equilibrium, :1213-1326
    run_operation; nobian/Simulation/Run.py:1399-1510): an EQUILIBRIUM stage with the creep elements only and every
    boundary value held at its first entry (:1156-1158), then ViscoplasticDesai is created, its hardening variable is
    initialised on the equilibrium stress (compute_initial_hardening(stress, Fvp_0=0.0), :1271-1274), it is added to the
    SAME material, and the OPERATION stage runs with compute_elastic_response=False (:1321-1324).
    Returns (case_eq, case_op); case_op["stage_elements"] are the elements to add between the stages."""
    return _staged(lambda els, n, is_eq: cavern_case(grid, elements=els, theta=theta, n_steps=n, ksp_type=ksp_type, rtol=rtol,
                                                     dt_hours=dt_eq_hours if is_eq else dt_op_hours),
                   creep, desai, n_eq, n_op)


def staged_triaxial_cases(grid, n_eq=2, n_op=3, creep=("kelvin", "dislocation"), desai="desai"):
    """The same two-stage workflow on the triaxial cube of BASELINE config 1 (what Simulators.py:1089-1326 runs for a
    GUI input file): equilibrium at the initial confining + axial load, then Desai (1_triaxial/main.py:78-89) with its
    hardening initialised on the equilibrium stress, then the axial load ramp."""
    return _staged(lambda els, n, is_eq: triaxial_case(grid, elements=els, n_steps=n), creep, desai, n_eq, n_op)


def add_operation_stage(case_op, eq, grid, verbose=False, outputs=None):
    """The lines between the two stages of the reference's scripts (Run.py:1494-1506, Simulators.py:1257-1276) and the
    operation stage's boundary conditions / simulator, on an equation that has run the equilibrium stage."""
    import safeincave_b200 as sf
    one = to.ones(grid.n_elems, dtype=to.float64)
    mat = eq.mat
    for e in case_op["stage_elements"]:
        desai = sf.ViscoplasticDesai(*[e[p] * one for p in sf.ViscoplasticDesai.param_names], e["alpha_0"] * one, "desai")
        stress_to = to.as_tensor(eq.sig.x.array.reshape((eq.n_elems, 3, 3)))      # Run.py:1500
        desai.compute_initial_hardening(stress_to, Fvp_0=0.0)
        mat.add_to_non_elastic(desai)
    eq.set_material(mat)
    bc = momBC.BcHandler(eq)
    for d in case_op["dirichlet"]:
        bc.add_boundary_condition(momBC.DirichletBC(d["boundary"], d["component"], d["values"], d["time_values"]))
    for nb in case_op["neumann"]:
        bc.add_boundary_condition(momBC.NeumannBC(nb["boundary"], nb["direction"], nb["density"], nb["ref_pos"],
                                                  nb["values"], nb["time_values"], g=nb["gravity"]))
    eq.set_boundary_conditions(bc)
    tc = sf.TimeController(dt=case_op["dt"], initial_time=0.0, final_time=case_op.get("t_final_run", case_op["t_final"]),
                           time_unit="second")
    return sf.Simulator_M(eq, tc, outputs or [], compute_elastic_response=False, verbose=verbose)


def per_cell(value, tm, n=None):
    """A case parameter as a per-cell float64 numpy array: a number (uniform) or a dict {region name: number} resolved
    through the mesh's cell tags (heterogeneous materials, e.g. salt / overburden of thermomechanics/2_cavern/main.py:47-97)."""
    n = tm.n_cells if n is None else n
    if isinstance(value, dict):
        out = np.zeros(n, dtype=np.float64)
        for name, v in value.items():
            out[tm.cell_tags == tm.names[3][name]] = float(v)
        return out
    return np.full(n, float(value), dtype=np.float64)


def overburden_tm_case(grid, n_steps=None, dt_days=0.5, ksp_type="cg", rtol=1e-12):
    """BASELINE config 4 (examples/thermomechanics/2_cavern/main.py on grids/cavern_overburden_coarse): salt + overburden
    with their own density / stiffness / viscosity / creep / thermal expansion (:47-97), rollers on the four sides of both
    layers and at the bottom (:127-138), free top, gas pressure on the cavern wall (:151-167), Spring + Thermoelastic +
    Kelvin + DislocationCreep + PressureSolutionCreep, theta = 0.5.  The thermal side (HeatDiffusion, :217-272) is in
    ``thermal``; the Thermoelastic element is added by the caller (``thermo_alpha`` per region)."""
    x = grid.mesh.geometry.x
    tm = grid.tetmesh
    z_top = float(x[:, 2].max())
    z_roof = float(x[np.unique(tm.tris[tm.tri_tags == grid.get_boundary_tag("Cavern")]), 2].max())
    ovb_cells = tm.cells[tm.cell_tags == tm.names[3]["Overburden"]]
    z_ovb = float(x[np.unique(ovb_cells), 2].min())                       # bottom of the overburden
    salt_density, ovb_density, gas_density, g = 2200.0, 2800.0, 0.082, -9.81
    p_roof = salt_density * abs(g) * (z_ovb - z_roof) + ovb_density * abs(g) * (z_top - z_ovb)
    t_final = 240 * day
    reg = lambda salt, ovb: {"Salt": salt, "Overburden": ovb}
    sides = [("West", 0), ("East", 0), ("South", 1), ("North", 1)]
    case = dict(
        name="overburden_tm", theta=0.5, dt=dt_days * day, t_final=t_final, time_unit="day",
        density=reg(salt_density, ovb_density), g=[0.0, 0.0, g],
        T=dict(surface=293.0, gradient=27.0 / 1000.0, z_surface=z_top),
        spring=dict(E=reg(102 * GPa, 180 * GPa), nu=0.3),
        elements=[dict(kind="kelvin", eta=reg(105e11, 105e21), E=10 * GPa, nu=0.32),
                  dict(kind="dislocation", A=reg(1.9e-20, 0.0), Q=51600.0, n=3.0),
                  dict(kind="pressure_solution", A=reg(1.29e-19, 0.0), d=0.01, Q=13184.0)],
        thermo_alpha=reg(44e-6, 0.0),
        dirichlet=[dict(boundary=f"{side}_{layer}", component=c, values=[0.0, 0.0], time_values=[0.0, t_final])
                   for side, c in sides for layer in ("salt", "ovb")]
        + [dict(boundary="Bottom", component=2, values=[0.0, 0.0], time_values=[0.0, t_final])],
        neumann=[dict(boundary="Top", direction=2, density=0.0, ref_pos=z_top, gravity=g, values=[0.0, 0.0],
                      time_values=[0.0, t_final]),
                 dict(boundary="Cavern", direction=2, density=gas_density, ref_pos=z_roof, gravity=g,
                      values=[0.8 * p_roof, 0.8 * p_roof, 0.2 * p_roof, 0.2 * p_roof, 0.8 * p_roof],
                      time_values=[0 * day, 20 * day, 40 * day, 60 * day, 80 * day])],
        thermal=dict(T0=293.0 + 0.027 * (z_top - x[:, 2]), t_final=t_final, rho=reg(salt_density, ovb_density), cp=850.0, k=7.0,
                     dirichlet=[dict(boundary="Top", values=[293.0, 293.0], time_values=[0.0, t_final])],
                     neumann=[dict(boundary="Bottom", values=[0.027, 0.027], time_values=[0.0, t_final])],
                     robin=[dict(boundary="Cavern", h=5.0, values=[293.0, 293.0], time_values=[0.0, t_final])]),
        ksp=dict(type=ksp_type, rtol=rtol), desai_initial_hardening=False,
    )
    if n_steps is not None:
        case["t_final_run"] = n_steps * case["dt"]
    return case


def cell_temperature(case, coords, cells):
    """Per-cell temperature (N,) float64 numpy."""
    T = case["T"]
    if isinstance(T, dict):
        zc = coords[cells][:, :, 2].mean(axis=1)
        return T["surface"] + T["gradient"] * (T["z_surface"] - zc)
    return np.full(cells.shape[0], float(T))


def build(case, grid, verbose=False, outputs=None, device="cuda", part=None, ctx=None):
    """Instantiate LinearMomentum + Material + BCs + Simulator_M for a case on the GPU, the way the
    reference's example scripts wire them.  ``grid`` may be a rank's local grid (then pass the
    Partition and DistContext; ``case`` must have been made from the GLOBAL grid)."""
    import safeincave_b200 as sf
    n = grid.n_elems

    class _PerCell:          # `value * one` for numbers and {region: value} dicts alike
        def __rmul__(self, v):
            return to.as_tensor(per_cell(v, grid.tetmesh, n))
    one = _PerCell()
    eq = sf.LinearMomentum(grid, theta=case["theta"], device=device)
    if ctx is not None:
        from . import distributed
        distributed.attach(eq, part, ctx)
    ksp = sf.PETSc.KSP().create(grid.mesh.comm)
    ksp.setType(case["ksp"]["type"])
    ksp.getPC().setType("asm")
    ksp.setTolerances(rtol=case["ksp"]["rtol"], max_it=case["ksp"].get("max_it", 100))
    eq.set_solver(ksp)
    mat = sf.Material(n)
    mat.set_density(case["density"] * one)
    mat.add_to_elastic(sf.Spring(case["spring"]["E"] * one, case["spring"]["nu"] * one, "spring"))
    for e in case["elements"]:
        k = e["kind"]
        if k == "kelvin":
            el = sf.Viscoelastic(e["eta"] * one, e["E"] * one, e["nu"] * one, "kelvin")
        elif k == "dislocation":
            el = sf.DislocationCreep(e["A"] * one, e["Q"] * one, e["n"] * one, "creep")
        elif k == "pressure_solution":
            el = sf.PressureSolutionCreep(e["A"] * one, e["d"] * one, e["Q"] * one, "pressure_solution")
        elif k == "desai":
            el = sf.ViscoplasticDesai(*[e[p] * one for p in sf.ViscoplasticDesai.param_names], e["alpha_0"] * one, "desai")
        elif k == "munson_dawson":
            el = sf.MunsonDawsonCreep(*[e[p] * one for p in sf.MunsonDawsonCreep.param_names], "munson_dawson")
        elif k in ("mohr_coulomb", "matsuoka_nakai"):
            cls = sf.MohrCoulombViscoplastic if k == "mohr_coulomb" else sf.MatsuokaNakaiViscoplastic
            el = cls(*[e[p] * one for p in cls.param_names], k)
        else:
            raise ValueError(k)
        mat.add_to_non_elastic(el)
    eq.set_material(mat)
    eq.build_body_force(case["g"])
    T = to.tensor(cell_temperature(case, grid.tetmesh.coords, grid.tetmesh.cells), dtype=to.float64)
    eq.set_T0(T)
    eq.set_T(T)
    bc = momBC.BcHandler(eq)
    for d in case["dirichlet"]:
        bc.add_boundary_condition(momBC.DirichletBC(d["boundary"], d["component"], d["values"], d["time_values"]))
    for nb in case["neumann"]:
        bc.add_boundary_condition(momBC.NeumannBC(nb["boundary"], nb["direction"], nb["density"], nb["ref_pos"],
                                                  nb["values"], nb["time_values"], g=nb["gravity"]))
    eq.set_boundary_conditions(bc)
    t_final = case.get("t_final_run", case["t_final"])
    tc = sf.TimeController(dt=case["dt"], initial_time=0.0, final_time=t_final, time_unit="second")
    sim = sf.Simulator_M(eq, tc, outputs or [], compute_elastic_response=True, verbose=verbose)
    return eq, sim
