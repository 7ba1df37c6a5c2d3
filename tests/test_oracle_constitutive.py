"""Pin the CPU oracle (oracle/constitutive.py) against the reference.

1. Known-answer vectors restated from the reference's own tests/test_material.py
   (values and tolerances cited by line).
2. Golden vectors produced by importing the unmodified reference in the build container
   (oracle/gen_golden.py).  Tolerance classes:
     EXACT  1e-12  everything that is not derived from a finite difference
     FD     2e-6   G, C_T, eps_rhs, Desai h/P/B/r: the reference's FD tangents carry
                   ~2e-8..1e-6 relative round-off of their own (hard part 1);
                   any other libm (glibc vs SLEEF vs ours) moves them by that much.
"""
import numpy as np
import pytest

from oracle import constitutive as oc
from tests import golden_replay as gr

EXACT = 1e-12
FD = 2e-6
FD_KEYS = ("tan:G", "tan:B", "tan:CT", "tan:eps_rhs", "tan:r", "tan:h", "tan:P", "tan:G_elem", "tan:B_elem")

STRESS = 1e6 * np.array([[[1., 4., 5.], [4., 2., 6.], [5., 6., 3.]]])
T298 = np.array([298.0])
DT, THETA = 7200.0, 0.5
PHI1, PHI2 = THETA * DT, (1 - THETA) * DT


def close(a, b, rtol, atol):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def test_spring_eps_e_kat():
    # reference tests/test_material.py:6-31 (E = 102 GPa, nu = 0.3; rtol 1e-6, atol 1e-9)
    E = 102e9 * np.ones(2)
    nu = 0.3 * np.ones(2)
    C_inv = np.linalg.inv(oc.iso_matrix(E, nu))
    sig = 1e6 * np.array([[[1., 4., 5.], [4., 2., 6.], [5., 6., 3.]],
                          [[6., 1., 2.], [1., 5., 3.], [2., 3., 4.]]])
    eps = oc.to_tensor(oc.spring_eps_e(C_inv, oc.to_voigt(sig)))
    true = np.array([[[-4.9020e-06, 5.0980e-05, 6.3725e-05], [5.0980e-05, 7.8431e-06, 7.6471e-05],
                      [6.3725e-05, 7.6471e-05, 2.0588e-05]],
                     [[3.2353e-05, 1.2745e-05, 2.5490e-05], [1.2745e-05, 1.9608e-05, 3.8235e-05],
                      [2.5490e-05, 3.8235e-05, 6.8627e-06]]])
    close(eps, true, 1e-4, 1e-9)   # the file prints 5 digits


def _run_elem(e, sig):
    e.tangent(sig, DT, THETA, T298)
    G = e.G.copy()
    e.eval_rate(sig, PHI1, T298)
    rate = e.rate.copy()
    e.predictor(PHI1, PHI2)
    eps_k = e.eps_k.copy()
    e.commit_strain(sig, np.zeros_like(sig), PHI2)
    return G, rate, eps_k, e.eps_old.copy()


def test_kelvin_kat():
    # reference tests/test_material.py:33-94
    e = oc.Kelvin(np.array([105e11]), np.array([10e9]), np.array([0.32]))
    G, rate, eps_k, eps_old = _run_elem(e, oc.to_voigt(STRESS))
    assert G[0, 0, 0] == pytest.approx(2.0666e-14, rel=1e-4)
    assert G[0, 0, 1] == pytest.approx(-5.8081e-15, rel=1e-4)
    assert G[0, 3, 3] == pytest.approx(2.6474e-14, rel=1e-4)
    true_rate = np.array([[[-8.3746e-09, 1.0590e-07, 1.3237e-07], [1.0590e-07, 1.8100e-08, 1.5884e-07],
                           [1.3237e-07, 1.5884e-07, 4.4574e-08]]])
    close(oc.to_tensor(rate), true_rate, 1e-4, 1e-10)
    true_k = np.array([[[-3.0148e-05, 3.8123e-04, 4.7653e-04], [3.8123e-04, 6.5158e-05, 5.7184e-04],
                        [4.7653e-04, 5.7184e-04, 1.6047e-04]]])
    close(oc.to_tensor(eps_k), true_k, 1e-4, 1e-8)
    true_old = np.array([[[-6.0297e-05, 7.6245e-04, 9.5307e-04], [7.6245e-04, 1.3032e-04, 1.1437e-03],
                          [9.5307e-04, 1.1437e-03, 3.2093e-04]]])
    close(oc.to_tensor(eps_old), true_old, 1e-4, 1e-7)


def test_dislocation_kat():
    # reference tests/test_material.py:96-152; G golden at 117-122 is NON-symmetric
    e = oc.Dislocation(np.array([1.9e-20]), np.array([51600.0]), np.array([3.0]))
    G, rate, eps_k, eps_old = _run_elem(e, oc.to_voigt(STRESS))
    true_G = np.array([[2.7650e-15, -1.3564e-15, -1.4086e-15, -8.3471e-16, -1.0434e-15, -1.2521e-15],
                       [-1.3564e-15, 2.7128e-15, -1.3564e-15, 0., 0., 0.],
                       [-1.4086e-15, -1.3564e-15, 2.7650e-15, 8.3471e-16, 1.0434e-15, 1.2521e-15],
                       [-2.0868e-16, 0., 2.0868e-16, 1.1477e-14, 4.1735e-15, 5.0083e-15],
                       [-2.6085e-16, 0., 2.6085e-16, 4.1735e-15, 1.3355e-14, 6.2603e-15],
                       [-3.1302e-16, 0., 3.1302e-16, 5.0083e-15, 6.2603e-15, 1.5651e-14]])
    close(G[0], true_G, 1e-4, 2e-19)
    true_rate = np.array([[[-4.0692e-09, 1.6277e-08, 2.0346e-08], [1.6277e-08, 0., 2.4415e-08],
                           [2.0346e-08, 2.4415e-08, 4.0692e-09]]])
    close(oc.to_tensor(rate), true_rate, 1e-4, 1e-10)
    true_k = np.array([[[-1.4649e-05, 5.8597e-05, 7.3246e-05], [5.8597e-05, 0., 8.7895e-05],
                        [7.3246e-05, 8.7895e-05, 1.4649e-05]]])
    close(oc.to_tensor(eps_k), true_k, 1e-4, 1e-8)
    true_old = np.array([[[-8.7519e-05, 4.0867e-04, 5.1084e-04], [4.0867e-04, 1.3643e-12, 6.1301e-04],
                          [5.1084e-04, 6.1301e-04, 8.7519e-05]]])
    close(oc.to_tensor(eps_old), true_old, 1e-4, 1e-4)


def test_pressure_solution_kat():
    # reference tests/test_material.py:154-213
    e = oc.PressureSolution(np.array([1.29e-15]), np.array([10e-3]), np.array([13184.0]))
    G, rate, eps_k, eps_old = _run_elem(e, oc.to_voigt(STRESS))
    assert G[0, 0, 0] == pytest.approx(1.4155e-14, rel=1e-4)
    assert G[0, 0, 1] == pytest.approx(-7.0777e-15, rel=1e-4)
    assert G[0, 3, 3] == pytest.approx(4.2466e-14, rel=1e-4)
    true_rate = np.array([[[-2.1233e-08, 8.4932e-08, 1.0617e-07], [8.4932e-08, 0., 1.2740e-07],
                           [1.0617e-07, 1.2740e-07, 2.1233e-08]]])
    close(oc.to_tensor(rate), true_rate, 1e-4, 1e-8)
    true_old = np.array([[[-1.5288e-04, 9.1727e-04, 1.1466e-03], [9.1727e-04, 7.1189e-12, 1.3759e-03],
                          [1.1466e-03, 1.3759e-03, 1.5288e-04]]])
    close(oc.to_tensor(eps_old), true_old, 1e-4, 1e-4)


def _desai_kat():
    from oracle.gen_golden import DESAI
    kw = {k: np.array([DESAI[k]]) for k in oc.DesaiParams.names}
    return oc.Desai(np.array([DESAI["alpha_0"]]), **kw)


def test_desai_kat_Fvp_alpha():
    # reference tests/test_material.py:215-285.  In this fork only Fvp (282) and alpha (283)
    # are still valid; G/rate/qsi goldens are stale and are pinned through the
    # imported reference instead (next test).
    e = _desai_kat()
    sig = oc.to_voigt(-1e7 * np.array([[[1., 0., 0.], [0., 1., 0.], [0., 0., 3.]]]))
    e.tangent(sig, DT, THETA, T298)
    e.eval_rate(sig, PHI1, T298)
    close(e.Fvp, [185.2260], 1e-3, 1e-4)
    close(e.alpha, [0.0022], 1e-3, 1e-4)


def test_kat_against_imported_reference():
    """Same inputs as the reference's tests, outputs of the imported reference itself."""
    g = dict(np.load(gr.GOLDEN_DIR + "/constitutive_kat.npz"))
    sig = oc.to_voigt(STRESS)
    for tag, e in (("kelvin", oc.Kelvin(np.array([105e11]), np.array([10e9]), np.array([0.32]))),
                   ("dislocation", oc.Dislocation(np.array([1.9e-20]), np.array([51600.0]), np.array([3.0]))),
                   ("pressure_solution", oc.PressureSolution(np.array([1.29e-15]), np.array([10e-3]), np.array([13184.0])))):
        G, rate, eps_k, eps_old = _run_elem(e, sig)
        assert gr.err(G, g[f"{tag}/G"]) < FD
        assert gr.err(rate, oc.to_voigt(g[f"{tag}/rate"])) < EXACT
        assert gr.err(eps_k, oc.to_voigt(g[f"{tag}/eps_k"])) < EXACT
        assert gr.err(eps_old, oc.to_voigt(g[f"{tag}/eps_old"])) < FD
    e = _desai_kat()
    sd = oc.to_voigt(-1e7 * np.array([[[1., 0., 0.], [0., 1., 0.], [0., 0., 3.]]]))
    G, rate, eps_k, eps_old = _run_elem(e, sd)
    assert gr.err(rate, oc.to_voigt(g["desai/rate"])) < EXACT
    assert gr.err(e.Fvp, g["desai/Fvp"]) < EXACT
    assert gr.err(e.qsi, g["desai/qsi"]) < 1e-9     # qsi left by the last P probe
    assert gr.err(G, g["desai/G"]) < 1e-4           # rank-one H/h term dominated by FD noise


GOLDENS = ["cfg1_kelvin_dc", "cfg2_dc_implicit", "cfg3_full", "desai_loaded", "thermo_het"]


@pytest.mark.parametrize("name", GOLDENS)
def test_oracle_matches_reference_goldens(name):
    g = gr.load(name)
    mat = gr.build_oracle_material(g)
    errs = gr.replay(g, mat, isolate=True)
    for k, v in errs.items():
        fd = k.startswith(FD_KEYS) or k.startswith("commit:eps_old")
        tol = FD if fd else EXACT
        if name == "desai_loaded" and k in ("tan:CT", "tan:eps_rhs", "tan:B", "tan:B_elem[desai]", "tan:r[desai]"):
            tol = 1e-3   # Desai alone: G ~ rank one, C_T = inv(C_inv + phi2 G) amplifies the FD noise
        if k.startswith(("tan:B", "tan:r")):
            tol = max(tol, 1e-3)   # r = alpha - a1/(..)^eta cancels to ~1e-4 of alpha; B = (r/h) Q
        assert v <= tol, f"{name}: {k} error {v:.3e} > {tol:.1e}"


# MunsonDawsonCreep, MohrCoulombViscoplastic, MatsuokaNakaiViscoplastic.  The reference's tests
# hold no vectors for them (tests/test_material.py stops at Desai): pinned by the imported reference only.
GOLDENS_EXT = ["md_alone", "md_implicit_het", "interlayer_mc", "interlayer_mn"]
# Munson-Dawson: n = 4.99 makes phi2*G two to three orders larger than C_inv, so C_T = inv(C_inv + phi2 G) and
# eps_rhs = ... - phi2 G:sigma amplify the reference's own FD round-off in G (2e-6) to 1e-4..1e-3; the zeta
# derivatives use a sqrt(eps)-sized step (MaterialProps.py:2013, 2257): h, P and everything downstream of the
# zeta increment carry 1e-9..1e-6.
EXT_TOL = {"tan:G": 5e-6, "tan:P": 5e-6, "tan:h": 5e-6, "tan:B": 1e-6, "tan:CT": 2e-3, "tan:eps_rhs": 5e-4, "commit:eps_old": 2e-5, "post:zeta": 1e-6, "commit:zeta_old": 1e-6,
           "post:rate": 1e-6, "post:F": 1e-6, "tan:r": 1e-6, "tan:eps_k": 1e-6}


@pytest.mark.parametrize("name", GOLDENS_EXT)
def test_oracle_matches_reference_goldens_extended_elements(name):
    g = gr.load(name)
    errs = gr.replay(g, gr.build_oracle_material(g), isolate=True)
    for k, v in errs.items():
        fd = k.startswith(FD_KEYS) or k.startswith("commit:eps_old")
        tol = FD if fd else EXACT
        if name.startswith("md_"):
            for key, t in EXT_TOL.items():
                if k.startswith(key):
                    tol = max(tol, t)
        else:
            tol = max(tol, 3e-6) if fd else tol          # Perzyna ramp ~ F^N: one more pow in the FD probes
        assert v <= tol, f"{name}: {k} error {v:.3e} > {tol:.1e}"
    # every quantity that does not pass through a finite difference is reproduced to round-off
    exact = {k: v for k, v in errs.items() if k.startswith(("init:", "post:sig", "post:Fvp", "commit:rate_old"))}
    assert exact and max(exact.values()) < EXACT


def test_interlayer_goldens_exercise_all_branches():
    """The interlayer stress states must sit on both sides of the yield surface and include the tension cut-off
    (at the initial state and hence in the first tangent phase; the overstress then relaxes within the step)."""
    for name, kind in (("interlayer_mc", "mohr_coulomb"), ("interlayer_mn", "matsuoka_nakai")):
        g = gr.load(name)
        Fvp = g["init/e1/Fvp"]
        I1 = -np.trace(g["sig0"], axis1=1, axis2=2) / 1e6
        tension = (-I1 / 3.0 - g[f"param/{kind}/sigma_t"]) > 0
        assert (Fvp > 0).sum() >= 12 and (Fvp <= 0).sum() >= 12 and tension.sum() >= 2
        rate = g["init/e1/eps_ne_rate"].reshape(len(Fvp), -1)
        active = np.abs(rate).max(axis=1) > 0
        assert active.sum() >= 10 and (active & ~tension).sum() >= 8          # shear-yielding cells flow
        assert not active[g[f"param/{kind}/mu_1"] == 0].any()                  # "salt" cells never do


def test_jacobi_eigenvalues_match_lapack():
    """The fixed six-sweep cyclic Jacobi iteration the kernels use (restated in eigvals_sym3_jacobi) against
    LAPACK, incl. repeated and well-separated eigenvalues."""
    rng = np.random.default_rng(0)
    s = np.concatenate([rng.standard_normal((3000, 6)) * 10,
                        np.array([[3., 3, 3, 0, 0, 0], [1, 2, 3, 0, 0, 0], [2, 2, 1, 1e-9, 0, 0], [5, 5, 5, 1, 1, 1],
                                  [1e3, 1e-3, 1, 0.1, 0.2, 0.3], [0, 0, 0, 0, 0, 0]])])
    a, b = oc.eigvals_sym3_jacobi(s), np.linalg.eigvalsh(oc.to_tensor(s))
    scale = np.maximum(np.abs(b).max(axis=1), 1e-300)
    assert (np.abs(a - b).max(axis=1) / scale).max() < 5e-15
    # and the Matsuoka-Nakai rate is insensitive to the choice
    g = gr.load("interlayer_mn")
    e = gr.build_oracle_material(g).elems[1]
    sig = gr.gold(g, "s0i0/sig")
    r1, f1 = e._rate(sig)
    oc.EIGEN = "jacobi"
    try:
        r2, f2 = e._rate(sig)
    finally:
        oc.EIGEN = "lapack"
    assert gr.err(r2, r1) < 1e-13 and gr.err(f2, f1) < 1e-13


def test_float32_user_parameters_deviation_is_bounded():
    """with float32 user tensors torch evaluates parts of the laws in float32.
    This build canonicalises parameters to float64 on entry; the deviation stays ~1e-6."""
    g = gr.load("cfg1_float32_params")
    mat = gr.build_oracle_material(g)
    errs = gr.replay(g, mat, isolate=True)
    assert max(errs.values()) < 5e-6


def test_libm_vs_sic_math_moves_G_by_fd_noise_only():
    """Swapping our exp/pow for numpy's (glibc/SVML) changes rates by < 1e-15 but G by ~1e-8:
    the floor no implementation can go below against the reference's FD tangent."""
    g = gr.load("cfg2_dc_implicit")
    sig = gr.gold(g, "s0i0/sig_k")
    T = g["T"]
    e = gr.build_oracle_material(g).elems[0]
    r1 = e._rate(sig, T)
    e.tangent(sig, float(g["dt"]), 0.0, T)
    G1 = e.G.copy()
    oc.USE_LIBM = True
    try:
        r2 = e._rate(sig, T)
        e.tangent(sig, float(g["dt"]), 0.0, T)
        G2 = e.G.copy()
    finally:
        oc.USE_LIBM = False
    assert gr.err(r1, r2) < 1e-15
    assert 0 < gr.err(G1, G2) < FD


def test_newton_error_counts_offdiagonals_twice():
    # Simulators.py:433-435 flattens the full 3x3 tensors
    a = np.random.default_rng(0).random((5, 6))
    b = np.random.default_rng(1).random((5, 6))
    ref = np.linalg.norm((oc.to_tensor(a) - oc.to_tensor(b)).ravel()) / np.linalg.norm(oc.to_tensor(b).ravel())
    assert oc.newton_error(a, b) == pytest.approx(ref, rel=1e-14)
