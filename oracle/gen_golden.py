"""Generate golden vectors for hot-path part (1) from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/gen_golden.py            # writes tests/golden/constitutive_*.npz

The reference's ``safeincave/MaterialProps.py`` is imported in place through
``oracle/refstub.py``.  ``LinearMomentum`` cannot be imported (dolfinx/petsc4py are
absent), so the few lines of constitutive glue it contributes are driven here exactly as
the reference's time loop drives them:

    Simulators.py:364-365       initial rates (phi1 = t0*theta, T7) + update_eps_ne_rate_old
    MomentumEquation.py:818-819 mat.compute_G_B ; mat.compute_CT
    MomentumEquation.py:887-889 eps_rhs = sum eps_ne_k + eps_th - phi2 (B + G:sigma_k)
    MomentumEquation.py:864     stress = CT : (eps_tot - eps_rhs)
    Simulators.py:422-425       increment_internal_variables ; compute_eps_ne_rate
    Simulators.py:509-517       update_internal_variables ; update_eps_ne_rate_old ; update_eps_ne_old

The total strain the linear solve would deliver is replaced by a deterministic synthetic
one (``eps_tot = C_inv : (sigma_k * (1 + load_step)) + eps_rhs``), identical in the tests.
Everything the reference objects hold after each phase is recorded.
"""
import os
import sys

import numpy as np
import torch as to

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.refstub import load_reference_material_props  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

DESAI = dict(mu_1=5.3665857009859815e-11, N_1=3.1, a_1=1.965018496922832e-05,
             eta=0.8275682807874163, n=3.0, beta_1=0.0048, beta=0.995, m=-0.5,
             gamma=0.095, sigma_t=5.0, alpha_0=0.0022)      # 1_triaxial/main.py:78-89


def make_inputs(N, seed):
    """Seeded compressive stress field -(10..30) MPa principal + scatter + small shear,
    T in [293, 320] K (SURVEY 8d 'kernel micro-benchmarks')."""
    g = np.random.default_rng(seed)
    sig = np.zeros((N, 3, 3))
    diag = -(10e6 + 20e6 * g.random((N, 3)))
    diag *= 1 + 0.1 * (g.random((N, 3)) - 0.5)
    shear = 2e6 * (g.random((N, 3)) - 0.5)
    for i in range(3):
        sig[:, i, i] = diag[:, i]
    for k, (i, j) in enumerate([(0, 1), (0, 2), (1, 2)]):
        sig[:, i, j] = shear[:, k]
        sig[:, j, i] = shear[:, k]
    T = 293 + 27 * g.random(N)
    noise = 0.02 * (g.random((N, 3, 3)) - 0.5)
    noise = 0.5 * (noise + noise.transpose(0, 2, 1))
    return sig, T, noise


MUNSON_DAWSON = dict(A=18.31 * (1e-6) ** 4.99 / (365 * 24 * 3600.0), Q=6356.0 * 8.32, n=4.99, K0=7.0e-7, c=9.02e-3,
                     m=3.0, alpha_w=-13.2, beta_w=-7.738, delta=0.58,
                     mu=20.425e9 / (2.0 * 1.25))      # nobian/Simulation/Run.py:1240-1276
INTERLAYER = dict(mu_1=1e-9, N_1=1.0, cohesion=4.0, friction_angle=float(np.radians(35.0)),
                  dilation_angle=float(np.radians(5.0)), sigma_t=1.0)   # run_interlayer.py:110-111, 1617-1623 (anhydrite)


def make_inputs_interlayer(N, seed):
    """Stress states around the Mohr-Coulomb / Matsuoka-Nakai surfaces of an anhydrite interlayer: low
    confinement -(0.5..6) MPa with one strongly compressed axis -(10..45) MPa (shear yield on roughly half
    of the cells), small shear, and a few cells in tension beyond sigma_t (tension cut-off)."""
    g = np.random.default_rng(seed)
    sig = np.zeros((N, 3, 3))
    diag = -(0.5e6 + 5.5e6 * g.random((N, 3)))
    axis = g.integers(0, 3, N)
    diag[np.arange(N), axis] = -(10e6 + 35e6 * g.random(N))
    n_t = max(2, N // 12)
    diag[:n_t] = 0.5e6 + 3e6 * g.random((n_t, 3))
    shear = 2e6 * (g.random((N, 3)) - 0.5)
    for i in range(3):
        sig[:, i, i] = diag[:, i]
    for k, (i, j) in enumerate([(0, 1), (0, 2), (1, 2)]):
        sig[:, i, j] = shear[:, k]
        sig[:, j, i] = shear[:, k]
    T = 293 + 27 * g.random(N)
    noise = 0.02 * (g.random((N, 3, 3)) - 0.5)
    noise = 0.5 * (noise + noise.transpose(0, 2, 1))
    return sig, T, noise


def build(mp, N, spec, dtype=to.float64, het=None):
    """spec: list of element kinds. het: optional (N,) multiplier making parameters
    heterogeneous."""
    one = to.ones(N, dtype=dtype)
    if het is not None:
        one = one * to.tensor(het, dtype=dtype)
    mat = mp.Material(N)
    mat.add_to_elastic(mp.Spring(102e9 * one, 0.3 * to.ones(N, dtype=dtype), "spring"))
    params = {"spring": dict(E=(102e9 * one).double().numpy(), nu=(0.3 * to.ones(N, dtype=dtype)).double().numpy())}
    for kind in spec:
        if kind == "kelvin":
            p = dict(eta=105e11 * one, E=10e9 * one, nu=0.32 * to.ones(N, dtype=dtype))
            mat.add_to_non_elastic(mp.Viscoelastic(p["eta"], p["E"], p["nu"], "kelvin"))
        elif kind == "dislocation":
            p = dict(A=1.9e-20 * one, Q=51600 * to.ones(N, dtype=dtype), n=3.0 * to.ones(N, dtype=dtype))
            mat.add_to_non_elastic(mp.DislocationCreep(p["A"], p["Q"], p["n"], "creep"))
        elif kind == "dislocation_n45":
            p = dict(A=1.1e-31 * one, Q=51600 * to.ones(N, dtype=dtype), n=4.5 * to.ones(N, dtype=dtype))
            mat.add_to_non_elastic(mp.DislocationCreep(p["A"], p["Q"], p["n"], "creep45"))
        elif kind == "pressure_solution":
            p = dict(A=1.29e-19 * one, d=0.01 * to.ones(N, dtype=dtype), Q=13184 * to.ones(N, dtype=dtype))
            mat.add_to_non_elastic(mp.PressureSolutionCreep(p["A"], p["d"], p["Q"], "ps"))
        elif kind == "desai":
            p = {k: v * to.ones(N, dtype=dtype) for k, v in DESAI.items()}
            mat.add_to_non_elastic(mp.ViscoplasticDesai(
                p["mu_1"], p["N_1"], p["a_1"], p["eta"], p["n"], p["beta_1"], p["beta"],
                p["m"], p["gamma"], p["sigma_t"], p["alpha_0"], "desai"))
        elif kind == "thermo":
            p = dict(alpha=44e-6 * to.ones(N, dtype=dtype))
            mat.add_to_thermoelastic(mp.Thermoelastic(p["alpha"], "thermo"))
        elif kind == "munson_dawson":
            p = {k: v * to.ones(N, dtype=dtype) for k, v in MUNSON_DAWSON.items()}
            p["A"] = p["A"] * one / one.mean() if het is not None else p["A"]
            mat.add_to_non_elastic(mp.MunsonDawsonCreep(**p, name="munson_dawson"))
        elif kind in ("mohr_coulomb", "matsuoka_nakai"):
            p = {k: v * to.ones(N, dtype=dtype) for k, v in INTERLAYER.items()}
            p["N_1"][::3] = 2.5                       # a fractional-free but non-unit rate exponent on a third of the cells
            p["mu_1"][1::7] = 0.0                     # "salt" cells that never yield (run_interlayer.py:1632-1638)
            cls = mp.MohrCoulombViscoplastic if kind == "mohr_coulomb" else mp.MatsuokaNakaiViscoplastic
            mat.add_to_non_elastic(cls(p["mu_1"], p["N_1"], p["cohesion"], p["friction_angle"], p["dilation_angle"],
                                       p["sigma_t"], kind))
        else:
            raise ValueError(kind)
        params[kind] = {k: v.double().numpy() for k, v in p.items()}
    return mat, params


def record_elems(mat, rec, tag):
    for i, e in enumerate(mat.elems_ne):
        pre = f"{tag}/e{i}"
        for name in ("eps_ne_rate", "eps_ne_rate_old", "eps_ne_old", "eps_ne_k", "G", "B"):
            rec[f"{pre}/{name}"] = getattr(e, name).double().numpy().copy()
        for name in ("alpha", "alpha_0", "Fvp", "qsi", "qsi_old", "r", "h", "P", "zeta", "zeta_old", "F"):
            if hasattr(e, name):
                rec[f"{pre}/{name}"] = getattr(e, name).double().numpy().copy()


def run_sequence(mp, name, N, seed, spec, dt, theta, n_steps=2, n_iters=3,
                 desai_init=False, load_step=0.03, dtype=to.float64, het=False, dT=0.0, inputs=None):
    sig0, T_np, noise = (inputs or make_inputs)(N, seed)
    het_arr = None
    if het:
        het_arr = 1 + 0.2 * (np.random.default_rng(seed + 1).random(N) - 0.5)
    mat, params = build(mp, N, spec, dtype=dtype, het=het_arr)
    dot = sys.modules["safeincave.Utils"].dotdot_torch
    T = to.tensor(T_np, dtype=to.float64)
    T0 = T - dT
    stress = to.tensor(sig0, dtype=to.float64)
    rec = {"sig0": sig0, "T": T_np, "T0": T0.numpy(), "noise": noise, "dt": dt, "theta": theta,
           "n_steps": n_steps, "n_iters": n_iters, "load_step": load_step,
           "desai_init": desai_init, "spec": np.array(spec)}
    for kind, p in params.items():
        for k, v in p.items():
            rec[f"param/{kind}/{k}"] = v
    rec["C"] = mat.C.numpy().copy()
    rec["C_inv"] = mat.C_inv.numpy().copy()
    if desai_init:
        for e in mat.elems_ne:
            if hasattr(e, "compute_initial_hardening"):
                e.compute_initial_hardening(stress, Fvp_0=0.0)
    # Simulators.py:364-365 (phi1 = t0*theta = 0)
    for e in mat.elems_ne:
        e.compute_eps_ne_rate(stress, 0.0 * theta, T, return_eps_ne=False)
    for e in mat.elems_ne:
        e.update_eps_ne_rate_old()
    record_elems(mat, rec, "init")
    phi1, phi2 = dt * theta, dt * (1 - theta)
    for step in range(n_steps):
        for it in range(n_iters):
            tag = f"s{step}i{it}"
            stress_k = stress.clone()
            # ---- tangent phase
            mat.compute_G_B(stress_k, dt, theta, T)
            mat.compute_CT(dt, theta)
            eps_ne_k = to.zeros((N, 3, 3), dtype=to.float64)
            for e in mat.elems_ne:
                e.compute_eps_ne_k(phi1, phi2)
                eps_ne_k += e.eps_ne_k
            eps_th = to.zeros((N, 3, 3), dtype=to.float64)
            for e in mat.elems_th:
                e.compute_eps_th(T - T0)
                eps_th += e.eps_th
            eps_rhs = eps_ne_k + eps_th - dt * (1 - theta) * (mat.B + dot(mat.G, stress_k))
            rec[f"{tag}/sig_k"] = stress_k.numpy().copy()
            rec[f"{tag}/G"] = mat.G.numpy().copy()
            rec[f"{tag}/B"] = mat.B.numpy().copy()
            rec[f"{tag}/CT"] = mat.CT.numpy().copy()
            rec[f"{tag}/eps_rhs"] = eps_rhs.numpy().copy()
            record_elems(mat, rec, tag + "/tan")
            # ---- synthetic "solve": strain of a slightly higher load level
            scale = 1 + load_step / (it + 1) * (1 + to.tensor(noise))
            eps_tot = dot(mat.C_inv, stress_k * scale) + eps_rhs
            rec[f"{tag}/eps_tot"] = eps_tot.numpy().copy()
            # ---- post phase
            stress = dot(mat.CT, eps_tot - eps_rhs)
            for e in mat.elems_ne:
                e.increment_internal_variables(stress, stress_k, dt)
            for e in mat.elems_ne:
                e.compute_eps_ne_rate(stress, dt * theta, T, return_eps_ne=False)
            rec[f"{tag}/sig"] = stress.numpy().copy()
            record_elems(mat, rec, tag + "/post")
        # ---- commit
        for e in mat.elems_ne:
            e.update_internal_variables()
        for e in mat.elems_ne:
            e.update_eps_ne_rate_old()
        for e in mat.elems_ne:
            e.update_eps_ne_old(stress, stress_k, dt * (1 - theta))
        record_elems(mat, rec, f"s{step}/commit")
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"constitutive_{name}.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, f"{os.path.getsize(path)/1024:.0f} KiB")


def kat_from_reference_tests(mp):
    """Re-run the reference's own tests/test_material.py inputs and store what the
    imported reference returns (the hard-coded 5-digit goldens of that file are restated
    as known-answer tests in tests/test_oracle_constitutive.py)."""
    rec = {}
    stress = 1e6 * to.tensor([[[1., 4., 5.], [4., 2., 6.], [5., 6., 3.]]], dtype=to.float64)
    T = 298 * to.ones(1, dtype=to.float64)
    theta, dt = 0.5, 7200.
    phi1, phi2 = theta * dt, (1 - theta) * dt
    zeros = to.zeros((1, 3, 3), dtype=to.float64)
    d = to.float64

    def run(tag, e, s=stress):
        e.compute_G_B(s, dt, theta, T)
        rec[f"{tag}/G"] = e.G.numpy().copy()
        e.compute_eps_ne_rate(s, phi1, T)
        rec[f"{tag}/rate"] = e.eps_ne_rate.numpy().copy()
        e.compute_eps_ne_k(phi1, phi2)
        rec[f"{tag}/eps_k"] = e.eps_ne_k.numpy().copy()
        e.update_eps_ne_old(s, zeros, phi2)
        rec[f"{tag}/eps_old"] = e.eps_ne_old.numpy().copy()
        for nm in ("Fvp", "alpha", "qsi"):
            if hasattr(e, nm):
                rec[f"{tag}/{nm}"] = getattr(e, nm).numpy().copy()

    run("kelvin", mp.Viscoelastic(105e11 * to.ones(1, dtype=d), 10e9 * to.ones(1, dtype=d), 0.32 * to.ones(1, dtype=d)))
    run("dislocation", mp.DislocationCreep(1.9e-20 * to.ones(1, dtype=d), 51600 * to.ones(1, dtype=d), 3.0 * to.ones(1, dtype=d)))
    run("pressure_solution", mp.PressureSolutionCreep(1.29e-15 * to.ones(1, dtype=d), 10e-3 * to.ones(1, dtype=d), 13184 * to.ones(1, dtype=d)))
    p = {k: v * to.ones(1, dtype=d) for k, v in DESAI.items()}
    s_desai = -1e7 * to.tensor([[[1., 0., 0.], [0., 1., 0.], [0., 0., 3.]]], dtype=to.float64)
    run("desai", mp.ViscoplasticDesai(p["mu_1"], p["N_1"], p["a_1"], p["eta"], p["n"], p["beta_1"], p["beta"],
                                      p["m"], p["gamma"], p["sigma_t"], p["alpha_0"]), s_desai)
    path = os.path.join(OUT, "constitutive_kat.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path)


def main_extended(mp):
    """MunsonDawsonCreep, MohrCoulombViscoplastic, MatsuokaNakaiViscoplastic."""
    H = 3600.0
    run_sequence(mp, "md_alone", 32, 21, ["munson_dawson"], 2 * H, 0.5, n_steps=2, n_iters=3, load_step=0.05)
    run_sequence(mp, "md_implicit_het", 32, 22, ["munson_dawson", "thermo"], 6 * H, 0.0, n_steps=2, n_iters=3,
                 load_step=0.08, het=True, dT=5.0)
    run_sequence(mp, "interlayer_mc", 48, 23, ["dislocation", "mohr_coulomb"], 1 * H, 0.5, load_step=0.04,
                 inputs=make_inputs_interlayer)
    run_sequence(mp, "interlayer_mn", 48, 24, ["dislocation", "matsuoka_nakai"], 1 * H, 0.0, load_step=0.04,
                 inputs=make_inputs_interlayer)


def main():
    mp = load_reference_material_props()
    to.manual_seed(0)
    if "--extended-only" in sys.argv:
        return main_extended(mp)
    kat_from_reference_tests(mp)
    H = 3600.0
    # config 1 (triaxial cube): Spring + Kelvin + DislocationCreep, theta 0.5, dt 0.5 h
    run_sequence(mp, "cfg1_kelvin_dc", 32, 1, ["kelvin", "dislocation"], 0.5 * H, 0.5)
    # config 2 (cavern_regular, fully implicit): Spring + DislocationCreep, theta 0, dt 2 h
    run_sequence(mp, "cfg2_dc_implicit", 32, 2, ["dislocation"], 2 * H, 0.0)
    # config 3: + PressureSolution + Desai (initial hardening at sig0, then loading)
    run_sequence(mp, "cfg3_full", 32, 3, ["kelvin", "dislocation", "pressure_solution", "desai"],
                 2 * H, 0.5, desai_init=True, load_step=0.15)
    # Desai alone after initial hardening, loaded by 20 %
    run_sequence(mp, "desai_loaded", 32, 4, ["desai"], 1 * H, 0.5, desai_init=True, load_step=0.2)
    # thermo-elastic strain + heterogeneous parameters + fractional stress exponent
    run_sequence(mp, "thermo_het", 32, 5, ["kelvin", "dislocation_n45", "pressure_solution", "thermo"],
                 12 * H, 0.5, het=True, dT=7.5)
    # float32 user parameters as in the examples -- documents the deviation
    run_sequence(mp, "cfg1_float32_params", 16, 6, ["kelvin", "dislocation"], 0.5 * H, 0.5, dtype=to.float32)
    main_extended(mp)


if __name__ == "__main__":
    main()
