"""Build the CPU oracle (OracleMaterial + OracleSimulatorM) from a safeincave_b200.cases dict."""
import numpy as np

from oracle import constitutive as oc
from oracle import fem
from safeincave_b200.cases import cell_temperature, per_cell


class _PerCell:              # `value * one` for numbers and {region: value} dicts alike (cases.per_cell)
    def __init__(self, tm, n):
        self.tm, self.n = tm, n

    def __rmul__(self, v):
        return per_cell(v, self.tm, self.n)


def oracle_material(case, n, tm=None):
    one = _PerCell(tm, n) if tm is not None else np.ones(n)
    mat = oc.OracleMaterial(n)
    mat.add_spring(case["spring"]["E"] * one, case["spring"]["nu"] * one)
    for e in case["elements"]:
        k = e["kind"]
        if k == "kelvin":
            mat.add(oc.Kelvin(e["eta"] * one, e["E"] * one, e["nu"] * one))
        elif k == "dislocation":
            mat.add(oc.Dislocation(e["A"] * one, e["Q"] * one, e["n"] * one))
        elif k == "pressure_solution":
            mat.add(oc.PressureSolution(e["A"] * one, e["d"] * one, e["Q"] * one))
        elif k == "desai":
            mat.add(oc.Desai(e["alpha_0"] * one, **{p: e[p] * one for p in oc.DesaiParams.names}))
        elif k == "munson_dawson":
            mat.add(oc.MunsonDawson(**{p: e[p] * one for p in oc.MunsonDawsonParams.names}))
        elif k in ("mohr_coulomb", "matsuoka_nakai"):
            cls = oc.MohrCoulomb if k == "mohr_coulomb" else oc.MatsuokaNakai
            mat.add(cls(*[e[p] * one for p in ("mu_1", "N_1", "cohesion", "friction_angle", "dilation_angle", "sigma_t")]))
    return mat


def oracle_simulator(case, tm):
    n = tm.n_cells
    mat = oracle_material(case, n, tm)
    T = cell_temperature(case, tm.coords, tm.cells)
    tag = lambda name: tm.names[2][name]
    dirichlet = [dict(tag=tag(d["boundary"]), component=d["component"], values=d["values"],
                      time_values=d["time_values"]) for d in case["dirichlet"]]
    neumann = [dict(tag=tag(b["boundary"]), direction=b["direction"], density=b["density"], ref_pos=b["ref_pos"],
                    gravity=b["gravity"], values=b["values"], time_values=b["time_values"]) for b in case["neumann"]]
    sim = fem.OracleSimulatorM(tm.coords, tm.cells, tm.tris, tm.tri_tags, mat, case["theta"], T, T,
                               per_cell(case["density"], tm, n), case["g"], dirichlet, neumann)
    if case.get("desai_initial_hardening"):
        def hook(m, sig):
            for e in m.elems:
                if e.kind == "desai":
                    e.initial_hardening(sig, 0.0)
        sim.after_initial_stress = hook
    return sim


def oracle_operation_stage(case_op, tm, osim_eq):
    """The operation stage of a staged run (Simulators.py:1213-1326; nobian/Simulation/Run.py:1494-1510) on the
    material, displacement and stress the oracle's equilibrium stage left behind: Desai is created, its hardening
    variable initialised on that stress (Fvp_0 = 0), and the new simulator starts WITHOUT an elastic response."""
    one = np.ones(tm.n_cells)
    for e in case_op["stage_elements"]:
        desai = oc.Desai(e["alpha_0"] * one, **{p: e[p] * one for p in oc.DesaiParams.names})
        desai.initial_hardening(osim_eq.sig, 0.0)
        osim_eq.mat.add(desai)
    sim = oracle_simulator(dict(case_op, elements=[]), tm)
    sim.mat = osim_eq.mat
    sim.compute_elastic_response = False
    sim.u, sim.sig = osim_eq.u.copy(), osim_eq.sig.copy()
    return sim


def oracle_staged_run(case_eq, case_op, tm):
    """Both stages; returns (equilibrium simulator, its history, operation simulator, its history)."""
    osim_eq = oracle_simulator(case_eq, tm)
    n_eq = int(round(case_eq["t_final_run"] / case_eq["dt"]))
    h_eq = osim_eq.run(0.0, [case_eq["dt"]] * n_eq)
    osim_op = oracle_operation_stage(case_op, tm, osim_eq)
    n_op = int(round(case_op["t_final_run"] / case_op["dt"]))
    h_op = osim_op.run(0.0, [case_op["dt"]] * n_op)
    return osim_eq, h_eq, osim_op, h_op
