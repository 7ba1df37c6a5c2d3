"""CPU oracle for hot-path part (1): the per-cell constitutive update.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module.
The product (``safeincave_b200``) never does.

This is a from-scratch numpy restatement (array-at-a-time, float64, Voigt-6 SoA-free
``(N, 6)`` arrays in the reference's order ``[xx, yy, zz, xy, xz, yz]`` with tensorial
shear) of the algorithm in the reference's ``safeincave/MaterialProps.py`` and of the
glue in ``safeincave/MomentumEquation.py``.  Every function cites the lines it follows.
Floating-point operation ORDER follows the reference's Python expressions (left to
right, one rounding per operation, no FMA), because the finite-difference tangents
amplify last-bit differences by ~5e8.

Pinning: ``tests/test_oracle_constitutive.py`` checks this module against
  * the known-answer vectors of the reference's own ``tests/test_material.py``
    (Spring 29-31, Viscoelastic 79-93, DislocationCreep 136-151,
    PressureSolutionCreep 197-212, Desai Fvp/alpha 282-283), and
  * golden vectors produced by importing the UNMODIFIED reference module in the build
    container (``oracle/gen_golden.py`` -> ``tests/golden/constitutive_*.npz``).

Elementary functions: exp/pow come from ``safeincave_b200/csrc/sic_math.h`` through the
ctypes shim ``oracle/sicmath_shim.c`` (``oracle/_build/libsicmath.so``) so that they are
bit-identical to the ones the CUDA kernels evaluate; ``use_libm=True`` switches to
numpy's own exp/power (used to show the difference is < 1 ulp).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

R_GAS = 8.32          # MaterialProps.py:915, 989 (not 8.314)
MPA = 1e6             # Utils.py:35
SQRT27 = float(np.sqrt(27))   # == 27**0.5
XX, YY, ZZ, XY, XZ, YZ = range(6)
VOIGT_PAIRS = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]

# --------------------------------------------------------------------------------------
# elementary functions (bit-identical to the CUDA side)
# --------------------------------------------------------------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_P = ctypes.POINTER(ctypes.c_double)
USE_LIBM = False


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libsicmath.so")
        if not os.path.isfile(path):
            raise RuntimeError(
                f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        _LIB = ctypes.CDLL(path)
    return _LIB


def f_exp(x):
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if USE_LIBM:
        with np.errstate(all="ignore"):
            return np.exp(x)
    out = np.empty_like(x)
    _lib().sic_exp_array(x.ctypes.data_as(_P), out.ctypes.data_as(_P), ctypes.c_long(x.size))
    return out


def f_pow(x, y):
    x, y = np.broadcast_arrays(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64))
    x = np.ascontiguousarray(x)
    y = np.ascontiguousarray(y)
    if USE_LIBM:
        with np.errstate(all="ignore"):
            return np.power(x, y)
    out = np.empty_like(x)
    _lib().sic_pow_array(x.ctypes.data_as(_P), y.ctypes.data_as(_P), out.ctypes.data_as(_P),
                         ctypes.c_long(x.size))
    return out


def f_log10(x):
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if USE_LIBM:
        with np.errstate(all="ignore"):
            return np.log10(x)
    out = np.empty_like(x)
    _lib().sic_log10_array(x.ctypes.data_as(_P), out.ctypes.data_as(_P), ctypes.c_long(x.size))
    return out


# --------------------------------------------------------------------------------------
# Voigt helpers
# --------------------------------------------------------------------------------------
def to_voigt(t33):
    """(N,3,3) -> (N,6) reading the UPPER entries, as every reference routine does
    (Utils.py:269-274, MaterialProps.py:941-946, 1214-1219)."""
    t33 = np.asarray(t33, dtype=np.float64)
    return np.stack([t33[:, i, j] for i, j in VOIGT_PAIRS], axis=1)


def to_tensor(v6):
    """(N,6) -> symmetric (N,3,3) (Utils.py:276-282)."""
    v6 = np.asarray(v6, dtype=np.float64)
    t = np.zeros((v6.shape[0], 3, 3))
    for k, (i, j) in enumerate(VOIGT_PAIRS):
        t[:, i, j] = v6[:, k]
        t[:, j, i] = v6[:, k]
    return t


def ddot(C, e):
    """sigma_v = C . eps_v, Utils.py:251-283 (``dotdot_torch``).  The six products of a
    row are accumulated left to right (the order the CUDA kernels use)."""
    r = C[:, :, 0] * e[:, None, 0]
    for k in range(1, 6):
        r = r + C[:, :, k] * e[:, None, k]
    return r


def iso_matrix(E, nu):
    """Isotropic stiffness in tensorial Voigt form, MaterialProps.py:459-487 (Spring) and
    825-833 (Viscoelastic C1): a0 = E/((1+nu)(1-2nu)); diag a0(1-nu); off-diag a0 nu;
    shear a0(1-2nu)."""
    E = np.asarray(E, dtype=np.float64)
    nu = np.asarray(nu, dtype=np.float64)
    n = E.shape[0]
    C = np.zeros((n, 6, 6))
    a0 = E / ((1 + nu) * (1 - 2 * nu))
    for i in range(3):
        C[:, i, i] = a0 * (1 - nu)
        C[:, i + 3, i + 3] = a0 * (1 - 2 * nu)
        for j in range(3):
            if i != j:
                C[:, i, j] = a0 * nu
    return C


def spring_eps_e(C_inv, sig):
    """Spring.compute_eps_e, MaterialProps.py:440-457."""
    return ddot(C_inv, sig)


# --------------------------------------------------------------------------------------
# rate laws  (all take sig as (N,6) in Pa and return the rate as (N,6))
# --------------------------------------------------------------------------------------
def _deviator(sig):
    """MaterialProps.py:948-952: only the diagonal is modified."""
    mean = (sig[:, XX] + sig[:, YY] + sig[:, ZZ]) / 3
    dev = sig.copy()
    dev[:, XX] = sig[:, XX] - mean
    dev[:, YY] = sig[:, YY] - mean
    dev[:, ZZ] = sig[:, ZZ] - mean
    return dev


def rate_dislocation(sig, T, A, Q, n):
    """DislocationCreep.compute_eps_ne_rate, MaterialProps.py:921-961."""
    sxx, syy, szz, sxy, sxz, syz = (sig[:, k] for k in range(6))
    dev = _deviator(sig)
    q = np.sqrt(0.5 * ((sxx - syy) ** 2 + (sxx - szz) ** 2 + (syy - szz) ** 2
                       + 6 * (sxy ** 2 + sxz ** 2 + syz ** 2)))
    with np.errstate(all="ignore"):
        A_bar = A * f_exp(-Q / R_GAS / T) * f_pow(q, n - 1)
    return A_bar[:, None] * dev


def rate_pressure_solution(sig, T, A, d, Q):
    """PressureSolutionCreep.compute_eps_ne_rate, MaterialProps.py:995-1034."""
    dev = _deviator(sig)
    with np.errstate(all="ignore"):
        A_bar = (A / (d * d * d) / T) * f_exp(-Q / R_GAS / T)
    return A_bar[:, None] * dev


def rate_kelvin(sig, G, C1, eps_old, rate_old, phi1):
    """Viscoelastic.compute_eps_ne_rate, MaterialProps.py:855: uses the G left behind by
    the last compute_G_B (zeros before the first one)."""
    return ddot(G, sig - ddot(C1, eps_old + phi1 * rate_old))


def kelvin_E(eta, C1, phi2):
    """Viscoelastic.compute_E, MaterialProps.py:882-885: (eta I + phi2 C1)^-1."""
    n = eta.shape[0]
    I = np.broadcast_to(np.eye(6), (n, 6, 6))
    return np.linalg.inv(eta[:, None, None] * I + phi2 * C1)


class DesaiParams:
    """Per-cell parameter arrays of ViscoplasticDesai (MaterialProps.py:1061-1092)."""
    names = ("mu_1", "N_1", "a_1", "eta", "n", "beta_1", "beta", "m", "gamma", "sigma_t")

    def __init__(self, **kw):
        for k in self.names:
            setattr(self, k, np.asarray(kw[k], dtype=np.float64))


def desai_invariants(sig, p):
    """extract_stress_components (1199-1220: sigma -> -sigma/MPa) followed by
    compute_stress_invariants (1160-1197)."""
    s = (-sig) / MPA
    sxx, syy, szz, sxy, sxz, syz = (s[:, k] for k in range(6))
    I1 = sxx + syy + szz
    I2 = sxx * syy + syy * szz + sxx * szz - sxy ** 2 - syz ** 2 - sxz ** 2
    I3 = (sxx * syy * szz + 2 * sxy * syz * sxz - szz * sxy ** 2 - sxx * syz ** 2
          - syy * sxz ** 2)
    J2 = (1 / 3) * I1 ** 2 - I2
    J3 = (2 / 27) * (I1 * I1 * I1) - (1 / 3) * I1 * I2 + I3
    J2_MIN = 1e-6
    with np.errstate(invalid="ignore"):
        low_J2 = J2 <= J2_MIN
    J2s = np.maximum(J2, J2_MIN)      # torch.clamp(min=): NaN stays NaN
    J2s = np.where(np.isnan(J2), J2, J2s)
    with np.errstate(all="ignore"):
        Sr = -(J3 * SQRT27) / (2 * f_pow(J2s, 1.5))
    Sr = np.where(low_J2, 0.0, Sr)
    I1s = I1 + p.sigma_t
    return s, I1, I2, I3, J2s, J3, Sr, I1s, low_J2


def _clamp_min(x, lo):
    """torch.clamp(x, min=lo): NaN propagates."""
    with np.errstate(invalid="ignore"):
        return np.where(x < lo, lo, x)


def desai_Fvp(alpha, I1s, J2, Sr, p):
    """compute_Fvp, MaterialProps.py:1222-1246."""
    with np.errstate(all="ignore"):
        F1 = alpha * f_pow(I1s, p.n) - p.gamma * I1s ** 2
        F2 = f_exp(p.beta_1 * I1s) - p.beta * Sr
        F2 = _clamp_min(F2, 1e-6)
        return J2 + F1 * f_pow(F2, p.m)


def rate_desai(sig, alpha, alpha_0, p):
    """ViscoplasticDesai.compute_eps_ne_rate, MaterialProps.py:1291-1429.
    Returns (rate (N,6), Fvp (N,))."""
    s, I1, I2, I3, J2, J3, Sr, I1s, low_J2 = desai_invariants(sig, p)
    sxx, syy, szz, sxy, sxz, syz = (s[:, k] for k in range(6))
    with np.errstate(all="ignore"):
        Fvp = desai_Fvp(alpha, I1s, J2, Sr, p)

        F1 = -alpha * f_pow(I1s, p.n) + p.gamma * I1s ** 2
        F2 = f_exp(p.beta_1 * I1s) - p.beta * Sr
        low_F2 = F2 < 1e-6
        F2 = _clamp_min(F2, 1e-6)

        F2m = f_pow(F2, p.m)
        F2m1 = f_pow(F2, p.m - 1)
        dF1_dI1 = 2 * p.gamma * I1s - p.n * alpha * f_pow(I1s, p.n - 1)
        dF2m_dI1 = p.beta_1 * p.m * f_exp(p.beta_1 * I1s) * F2m1
        dF_dI1 = -(dF1_dI1 * F2m + F1 * dF2m_dI1)

        J2_15 = f_pow(J2, 1.5)
        dF2_dJ2 = -(3 * p.beta * J3 * SQRT27) / (4 * f_pow(J2, 5 / 2))
        dF_dJ2 = 1 - F1 * p.m * F2m1 * dF2_dJ2
        dF_dJ3 = -p.m * F1 * p.beta * SQRT27 * F2m1 / (2 * J2_15)

        dI1 = (1.0, 1.0, 1.0, 0.0, 0.0, 0.0)
        dI2 = (syy + szz, sxx + szz, sxx + syy, -2 * sxy, -2 * sxz, -2 * syz)
        dI3 = (syy * szz - syz ** 2, sxx * szz - sxz ** 2, sxx * syy - sxy ** 2,
               2 * (sxz * syz - szz * sxy), 2 * (sxy * syz - syy * sxz),
               2 * (sxz * sxy - sxx * syz))
        dJ2_dI1 = (2 / 3) * I1
        dJ2_dI2 = -1.0
        dJ3_dI1 = (2 / 9) * I1 ** 2 - (1 / 3) * I2
        dJ3_dI2 = -(1 / 3) * I1
        dJ3_dI3 = 1.0

        dQ = np.zeros_like(sig)
        for k in range(6):
            dJ2_dS = dJ2_dI1 * dI1[k] + dJ2_dI2 * dI2[k]
            dJ3_dS = dJ3_dI1 * dI1[k] + dJ3_dI2 * dI2[k] + dJ3_dI3 * dI3[k]
            dQ[:, k] = dF_dI1 * dI1[k] + dF_dJ2 * dJ2_dS + dF_dJ3 * dJ3_dS

        softened = alpha <= 0.01 * alpha_0
        off = low_J2 | low_F2 | softened
        dQ[off, :] = 0.0

        ramp = Fvp > 0
        lam = np.zeros_like(Fvp)
        if ramp.any():
            lam[ramp] = p.mu_1[ramp] * f_pow(Fvp[ramp] / 1.0, p.N_1[ramp])
        rate = -dQ * lam[:, None]
    return rate, Fvp


def frob_norm_sym(v):
    """to.sum(rate**2, axis=(-2,-1))**0.5 over the 9 entries of the symmetric tensor
    (MaterialProps.py:1116); summation order fixed as diag-sum + 2*offdiag-sum."""
    d = v[:, XX] ** 2 + v[:, YY] ** 2 + v[:, ZZ] ** 2
    o = v[:, XY] ** 2 + v[:, XZ] ** 2 + v[:, YZ] ** 2
    return np.sqrt(d + 2 * o)


def desai_residue(rate, alpha, qsi_old, alpha_0, dt, p):
    """compute_residue, MaterialProps.py:1094-1117.  Returns (r, qsi)."""
    with np.errstate(all="ignore"):
        qsi = qsi_old + frob_norm_sym(rate) * dt
        r = alpha - p.a_1 / f_pow(f_pow(p.a_1 / alpha_0, 1 / p.eta) + qsi, p.eta)
    return r, qsi


def ddot_sym(a, b):
    """einsum('bij,bij->b') of two symmetric tensors held as Voigt-6 (MaterialProps.py:1150)."""
    d = a[:, XX] * b[:, XX] + a[:, YY] * b[:, YY] + a[:, ZZ] * b[:, ZZ]
    o = a[:, XY] * b[:, XY] + a[:, XZ] * b[:, XZ] + a[:, YZ] * b[:, YZ]
    return d + 2 * o


# --------------------------------------------------------------------------------------
# finite-difference tangent, NonElasticElement.compute_E  (MaterialProps.py:640-675)
# --------------------------------------------------------------------------------------
def fd_tangent(rate_fn, sig):
    """Central FD, eps = 1e-2 Pa, running-copy perturbation (+=, -=, -=, += on ONE
    buffer), shear columns doubled (T3)."""
    EPS = 1e-2
    n = sig.shape[0]
    E = np.zeros((n, 6, 6))
    s = sig.copy()
    for k in range(6):
        phi = 1.0 if k < 3 else 2.0
        s[:, k] += EPS
        ra = rate_fn(s)
        s[:, k] -= EPS
        s[:, k] -= EPS
        rb = rate_fn(s)
        s[:, k] += EPS
        E[:, :, k] = phi * (ra - rb) / (2 * EPS)
    return E


# --------------------------------------------------------------------------------------
# elements + material (state containers; mirror the reference objects' data flow)
# --------------------------------------------------------------------------------------
class _Element:
    kind = "?"

    def __init__(self, n):
        self.n = n
        self.rate = np.zeros((n, 6))        # eps_ne_rate
        self.rate_old = np.zeros((n, 6))    # eps_ne_rate_old
        self.eps_old = np.zeros((n, 6))     # eps_ne_old
        self.eps_k = np.zeros((n, 6))       # eps_ne_k
        self.G = np.zeros((n, 6, 6))
        self.B = np.zeros((n, 6))

    # NonElasticElement.compute_eps_ne_k, MaterialProps.py:586-605
    def predictor(self, phi1, phi2):
        self.eps_k = self.eps_old + phi1 * self.rate_old + phi2 * self.rate

    # NonElasticElement.update_eps_ne_old, MaterialProps.py:607-628
    def commit_strain(self, sig, sig_k, phi2):
        self.eps_old = self.eps_k + phi2 * ddot(self.G, sig - sig_k) - phi2 * self.B

    # NonElasticElement.update_eps_ne_rate_old, MaterialProps.py:630-638
    def commit_rate(self):
        self.rate_old = self.rate.copy()

    def increment_isv(self, sig, sig_k, dt):
        pass

    def commit_isv(self):
        pass

    def snapshot(self):
        return {k: getattr(self, k).copy() for k in ("rate", "rate_old", "eps_old", "eps_k")}

    def restore(self, snap):
        for k, v in snap.items():
            setattr(self, k, v.copy())


class Kelvin(_Element):
    kind = "kelvin"

    def __init__(self, eta, E, nu):
        super().__init__(len(eta))
        self.eta = np.asarray(eta, dtype=np.float64)
        self.C1 = iso_matrix(E, nu)

    def tangent(self, sig, dt, theta, T):      # compute_G_B, 707-728 with compute_E 861-885
        self.B = np.zeros((self.n, 6))
        self.G = kelvin_E(self.eta, self.C1, dt * (1 - theta))

    def eval_rate(self, sig, phi1, T):
        self.rate = rate_kelvin(sig, self.G, self.C1, self.eps_old, self.rate_old, phi1)


class Dislocation(_Element):
    kind = "dislocation"

    def __init__(self, A, Q, n):
        super().__init__(len(A))
        self.A, self.Q, self.nexp = (np.asarray(x, dtype=np.float64) for x in (A, Q, n))

    def _rate(self, sig, T):
        return rate_dislocation(sig, T, self.A, self.Q, self.nexp)

    def tangent(self, sig, dt, theta, T):
        self.B = np.zeros((self.n, 6))
        self.G = fd_tangent(lambda s: self._rate(s, T), sig)

    def eval_rate(self, sig, phi1, T):
        self.rate = self._rate(sig, T)


class PressureSolution(_Element):
    kind = "pressure_solution"

    def __init__(self, A, d, Q):
        super().__init__(len(A))
        self.A, self.d, self.Q = (np.asarray(x, dtype=np.float64) for x in (A, d, Q))

    def _rate(self, sig, T):
        return rate_pressure_solution(sig, T, self.A, self.d, self.Q)

    def tangent(self, sig, dt, theta, T):
        self.B = np.zeros((self.n, 6))
        self.G = fd_tangent(lambda s: self._rate(s, T), sig)

    def eval_rate(self, sig, phi1, T):
        self.rate = self._rate(sig, T)


class Desai(_Element):
    kind = "desai"

    def __init__(self, alpha_0, **params):
        alpha_0 = np.asarray(alpha_0, dtype=np.float64)
        super().__init__(len(alpha_0))
        self.p = DesaiParams(**params)
        self.alpha_0 = alpha_0.copy()
        self.alpha = alpha_0.copy()                      # :1089
        self.Fvp = np.zeros(self.n)
        self.qsi = np.zeros(self.n)
        self.qsi_old = np.zeros(self.n)
        self.r = np.zeros(self.n)
        self.h = np.ones(self.n)
        self.P = np.zeros((self.n, 6))
        self.h_small = np.zeros(self.n, dtype=bool)

    def initial_hardening(self, sig, Fvp_0=0.0):
        """compute_initial_hardening, MaterialProps.py:1248-1288."""
        p = self.p
        _, I1, I2, I3, J2, J3, Sr, I1s, _ = desai_invariants(sig, p)
        with np.errstate(all="ignore"):
            F2 = f_exp(p.beta_1 * I1s) - p.beta * Sr
            F2 = _clamp_min(F2, 1e-6)
            a0 = p.gamma * f_pow(I1s, 2 - p.n) + (Fvp_0 - J2) * f_pow(I1s, -p.n) * f_pow(F2, -p.m)
        self.n_disabled = int((a0 <= 1e-6).sum())
        self.alpha_0 = _clamp_min(a0, 1e-6)
        self.alpha = self.alpha_0.copy()
        _, I1, I2, I3, J2, J3, Sr, I1s, _ = desai_invariants(sig, p)
        self.Fvp = desai_Fvp(self.alpha, I1s, J2, Sr, p)

    def _rate(self, sig, alpha):
        return rate_desai(sig, alpha, self.alpha_0, self.p)[0]

    def tangent(self, sig, dt, theta, T):
        """compute_G_B (707-728) = compute_B_and_H_over_h (1432-1500) then compute_E
        (640-675) and compute_H (1503-1562)."""
        p = self.p
        eps_alpha = 0.0001 * self.alpha
        alpha_eps = self.alpha + eps_alpha
        rate_eps = self._rate(sig, alpha_eps)
        self.r, self.qsi = desai_residue(self.rate, self.alpha, self.qsi_old, self.alpha_0, dt, p)
        r_eps, self.qsi = desai_residue(rate_eps, alpha_eps, self.qsi_old, self.alpha_0, dt, p)
        with np.errstate(all="ignore"):
            self.h = (r_eps - self.r) / eps_alpha
            Q = (rate_eps - self.rate) / eps_alpha[:, None]
            self.h_small = np.abs(self.h) < 1e-6
            self.h = np.where(self.h_small, 1.0, self.h)
            B = (self.r / self.h)[:, None] * Q
        EPS_S = 1e-1
        self.P = np.zeros((self.n, 6))
        s = sig.copy()
        for k in range(6):
            s[:, k] += EPS_S
            rate_p = self._rate(s, self.alpha)
            r_p, self.qsi = desai_residue(rate_p, self.alpha, self.qsi_old, self.alpha_0, dt, p)
            self.P[:, k] = (r_p - self.r) / EPS_S
            s[:, k] -= EPS_S
        H = np.zeros((self.n, 6, 6))
        for i in range(6):
            for j in range(6):
                if j < 3:
                    H[:, i, j] = Q[:, i] * self.P[:, j]
                else:
                    H[:, i, j] = 2 * Q[:, i] * self.P[:, j]
        with np.errstate(all="ignore"):
            H_over_h = H / self.h[:, None, None]
        B[self.h_small] = 0.0
        H_over_h[self.h_small] = 0.0
        self.P[self.h_small] = 0.0
        self.B = B
        E = fd_tangent(lambda x: self._rate(x, self.alpha), sig)
        self.G = E - H_over_h

    def eval_rate(self, sig, phi1, T):
        self.rate, self.Fvp = rate_desai(sig, self.alpha, self.alpha_0, self.p)

    def increment_isv(self, sig, sig_k, dt):
        """increment_internal_variables, MaterialProps.py:1129-1158."""
        with np.errstate(all="ignore"):
            d_alpha = -(self.r + ddot_sym(self.P, sig - sig_k)) / self.h
        d_alpha = np.where(self.h_small, 0.0, d_alpha)
        self.alpha = self.alpha + d_alpha
        self.alpha = _clamp_min(self.alpha, 1e-10)

    def commit_isv(self):
        """update_internal_variables, MaterialProps.py:1119-1127."""
        self.qsi_old = self.qsi.copy()

    def snapshot(self):
        snap = super().snapshot()
        for k in ("alpha", "qsi", "qsi_old", "Fvp"):
            snap[k] = getattr(self, k).copy()
        return snap



# --------------------------------------------------------------------------------------
# MunsonDawsonCreep, MaterialProps.py:1971-2346
# --------------------------------------------------------------------------------------
class MunsonDawsonParams:
    names = ("A", "Q", "n", "K0", "c", "m", "alpha_w", "beta_w", "delta", "mu")

    def __init__(self, **kw):
        for k in self.names:
            setattr(self, k, np.asarray(kw[k], dtype=np.float64))


def _clamp(x, lo, hi):
    """torch.clamp(x, min=lo, max=hi): NaN propagates."""
    with np.errstate(invalid="ignore"):
        return np.where(x < lo, lo, np.where(x > hi, hi, x))


def md_fields(sig, T, zeta, p):
    """_compute_md_fields, MaterialProps.py:2104-2167.  Returns (s_dev (N,6), sigma_safe, epsdot_ss, eps_t_star, F)."""
    sxx, syy, szz, sxy, sxz, syz = (sig[:, k] for k in range(6))
    dev = _deviator(sig)
    sigma = np.sqrt(0.5 * ((sxx - syy) ** 2 + (sxx - szz) ** 2 + (syy - szz) ** 2
                           + 6.0 * (sxy ** 2 + sxz ** 2 + syz ** 2)))
    sigma_safe = _clamp_min(sigma, 1.0)
    mu_safe = _clamp_min(p.mu, 1.0)
    with np.errstate(all="ignore"):
        epsdot_ss = p.A * f_exp(-p.Q / (R_GAS * T)) * f_pow(sigma_safe, p.n)
        ratio = _clamp_min(sigma_safe / mu_safe, 1e-30)
        ets = p.K0 * f_exp(p.c * T) * f_pow(ratio, p.m)
        ets = _clamp_min(ets, 1e-50)
        Delta = p.alpha_w + p.beta_w * f_log10(ratio)
        r_arg = 1.0 - (zeta / ets)
        r2 = r_arg * r_arg
        hard = zeta <= ets
        arg = np.where(hard, _clamp(Delta * r2, -50.0, 50.0), _clamp(-p.delta * r2, -50.0, 50.0))
        F = f_exp(arg)
    return dev, sigma_safe, epsdot_ss, ets, F


def rate_munson_dawson(sig, T, zeta, p):
    """compute_eps_ne_rate, MaterialProps.py:2187-2229.  Returns (rate, F, eps_t_star)."""
    dev, sigma_safe, epsdot_ss, ets, F = md_fields(sig, T, zeta, p)
    scalar_rate = F * epsdot_ss
    flow = (1.5 / sigma_safe)[:, None] * dev
    return flow * scalar_rate[:, None], F, ets


def md_residue(sig, T, zeta, zeta_old, dt, p):
    """compute_residue, MaterialProps.py:2169-2181."""
    _, _, epsdot_ss, _, F = md_fields(sig, T, zeta, p)
    return zeta - zeta_old - (F - 1.0) * epsdot_ss * dt


class MunsonDawson(_Element):
    kind = "munson_dawson"
    SQRT_EPS = 1.4901161193847656e-8          # MaterialProps.py:2013

    def __init__(self, **params):
        self.p = MunsonDawsonParams(**params)
        super().__init__(len(self.p.A))
        n = self.n
        self.zeta, self.zeta_old = np.zeros(n), np.zeros(n)
        self.F, self.eps_t_star = np.ones(n), np.ones(n)
        self.r, self.h, self.P = np.zeros(n), np.ones(n), np.zeros((n, 6))
        self.h_small = np.zeros(n, dtype=bool)

    def tangent(self, sig, dt, theta, T):
        """compute_G_B (707-728) = compute_B_and_H_over_h (2235-2313), compute_E (640-675), _compute_H (2315-2346)."""
        p = self.p
        _, _, _, ets_now, _ = md_fields(sig, T, self.zeta, p)
        zeta_scale = _clamp_min(np.abs(self.zeta) + ets_now, 1e-30)
        eps_zeta = self.SQRT_EPS * zeta_scale
        self.r = md_residue(sig, T, self.zeta, self.zeta_old, dt, p)
        zeta_eps = self.zeta + eps_zeta
        r_zeta = md_residue(sig, T, zeta_eps, self.zeta_old, dt, p)
        with np.errstate(all="ignore"):
            self.h = (r_zeta - self.r) / eps_zeta
            rate_ref = rate_munson_dawson(sig, T, self.zeta, p)[0]
            rate_zeta = rate_munson_dawson(sig, T, zeta_eps, p)[0]
            Q = (rate_zeta - rate_ref) / eps_zeta[:, None]
            self.h_small = np.abs(self.h) < 1e-12
            self.h = np.where(self.h_small, 1.0, self.h)
            B = (self.r / self.h)[:, None] * Q
        EPS_S = 1e-1
        self.P = np.zeros((self.n, 6))
        s = sig.copy()
        for k in range(6):
            s[:, k] += EPS_S
            r_sig = md_residue(s, T, self.zeta, self.zeta_old, dt, p)
            self.P[:, k] = (r_sig - self.r) / EPS_S
            s[:, k] -= EPS_S
        H = np.zeros((self.n, 6, 6))
        for i in range(6):
            for j in range(6):
                H[:, i, j] = Q[:, i] * self.P[:, j] if j < 3 else 2 * Q[:, i] * self.P[:, j]
        with np.errstate(all="ignore"):
            H_over_h = H / self.h[:, None, None]
        B[self.h_small] = 0.0
        H_over_h[self.h_small] = 0.0
        self.P[self.h_small] = 0.0
        self.B = B
        E = fd_tangent(lambda x: rate_munson_dawson(x, T, self.zeta, p)[0], sig)
        self.G = E - H_over_h

    def eval_rate(self, sig, phi1, T):
        self.rate, self.F, self.eps_t_star = rate_munson_dawson(sig, T, self.zeta, self.p)

    def increment_isv(self, sig, sig_k, dt):
        """increment_internal_variables, MaterialProps.py:2081-2102."""
        with np.errstate(all="ignore"):
            d = -(self.r + ddot_sym(self.P, sig - sig_k)) / self.h
        d = np.where(self.h_small, 0.0, d)
        self.zeta = _clamp_min(self.zeta + d, 0.0)

    def commit_isv(self):
        """update_internal_variables, MaterialProps.py:2077-2079."""
        self.zeta_old = self.zeta.copy()

    def snapshot(self):
        snap = super().snapshot()
        for k in ("zeta", "zeta_old"):
            snap[k] = getattr(self, k).copy()
        return snap


# --------------------------------------------------------------------------------------
# MohrCoulombViscoplastic (1565-1746) and MatsuokaNakaiViscoplastic (1749-1968)
# --------------------------------------------------------------------------------------
def _dp_flow(s, I1, alpha_Q, is_tension):
    """DP-based flow direction shared by both models (1705-1731 / 1927-1953): dQ/dsigma in the
    compression-positive MPa convention, tension cells replaced by -1/3 on the diagonal."""
    sxx, syy, szz, sxy, sxz, syz = (s[:, k] for k in range(6))
    I2 = sxx * syy + syy * szz + sxx * szz - sxy ** 2 - syz ** 2 - sxz ** 2
    J2 = _clamp_min((1.0 / 3.0) * I1 ** 2 - I2, 1e-20)
    sqrt_J2 = np.sqrt(J2)
    inv = 1.0 / (2.0 * sqrt_J2)
    dJ2 = ((2.0 / 3.0) * I1 - (syy + szz), (2.0 / 3.0) * I1 - (sxx + szz), (2.0 / 3.0) * I1 - (sxx + syy),
           2.0 * sxy, 2.0 * sxz, 2.0 * syz)
    dQ = np.zeros_like(s)
    for k in range(6):
        dQ[:, k] = inv * dJ2[k] - alpha_Q if k < 3 else inv * dJ2[k]
    dQ[is_tension, :] = 0.0
    dQ[is_tension, :3] = -1.0 / 3.0
    return dQ, sqrt_J2


def _perzyna(Fvp, mu_1, N_1):
    lam = np.zeros_like(Fvp)
    with np.errstate(invalid="ignore"):
        ramp = Fvp > 0
    if ramp.any():
        with np.errstate(all="ignore"):
            lam[ramp] = mu_1[ramp] * f_pow(Fvp[ramp] / 1.0, N_1[ramp])
    return lam


def rate_mohr_coulomb(sig, mu_1, N_1, alpha_F, k_F, alpha_Q, sigma_t):
    """MohrCoulombViscoplastic.compute_eps_ne_rate, MaterialProps.py:1652-1746.  Returns (rate, Fvp)."""
    s = (-sig) / MPA
    I1 = s[:, XX] + s[:, YY] + s[:, ZZ]
    F_tension = -I1 / 3.0 - sigma_t
    # the flow direction needs sqrt(J2) as well; the yield function uses the same value
    dQ_shear_only, sqrt_J2 = _dp_flow(s, I1, alpha_Q, np.zeros(len(I1), dtype=bool))
    F_shear = sqrt_J2 - alpha_F * I1 - k_F
    Fvp = np.maximum(F_shear, F_tension)
    is_tension = F_tension > F_shear
    dQ = dQ_shear_only
    dQ[is_tension, :] = 0.0
    dQ[is_tension, :3] = -1.0 / 3.0
    lam = _perzyna(Fvp, mu_1, N_1)
    return -dQ * lam[:, None], Fvp


EIGEN = "lapack"     # "jacobi": the fixed six-sweep cyclic Jacobi iteration the CUDA kernels run (bit-identical)


def _jacobi_rotate(app, aqq, apq, arp, arq):
    nz = apq != 0.0
    with np.errstate(all="ignore"):
        theta = np.where(nz, (aqq - app) / np.where(nz, 2.0 * apq, 1.0), 0.0)
        t = 1.0 / (np.abs(theta) + np.sqrt(theta * theta + 1.0))
    t = np.where(theta < 0.0, -t, t)
    t = np.where(nz, t, 0.0)
    c = 1.0 / np.sqrt(t * t + 1.0)
    sn = t * c
    return app - t * apq, aqq + t * apq, np.zeros_like(apq), c * arp - sn * arq, sn * arp + c * arq


def eigvals_sym3_jacobi(s):
    """Ascending eigenvalues of symmetric 3x3 tensors given as Voigt-6: restatement, operation for
    operation, of eigvals_sym3 in safeincave_b200/csrc/constitutive.cuh."""
    a00, a11, a22, a01, a02, a12 = (s[:, k].copy() for k in range(6))
    for _ in range(6):
        a00, a11, a01, a02, a12 = _jacobi_rotate(a00, a11, a01, a02, a12)
        a00, a22, a02, a01, a12 = _jacobi_rotate(a00, a22, a02, a01, a12)
        a11, a22, a12, a01, a02 = _jacobi_rotate(a11, a22, a12, a01, a02)
    x, y, z = a00, a11, a22
    x, y = np.where(y < x, y, x), np.where(y < x, x, y)
    y, z = np.where(z < y, z, y), np.where(z < y, y, z)
    x, y = np.where(y < x, y, x), np.where(y < x, x, y)
    return np.stack([x, y, z], axis=1)


def rate_matsuoka_nakai(sig, mu_1, N_1, k_nfc, shift, alpha_Q, sigma_t):
    """MatsuokaNakaiViscoplastic.compute_eps_ne_rate, MaterialProps.py:1835-1968.  Returns (rate, Fvp)."""
    s = (-sig) / MPA
    if EIGEN == "jacobi":
        eig = eigvals_sym3_jacobi(s)
    else:
        eig = np.linalg.eigvalsh(to_tensor(s))             # ascending (torch.linalg.eigvalsh, :1882)
    sig3, sig2, sig1 = eig[:, 0], eig[:, 1], eig[:, 2]
    s1, s2, s3 = sig1 + shift, sig2 + shift, sig3 + shift
    d12, d23, d31 = _clamp_min(s1 + s2, 1e-20), _clamp_min(s2 + s3, 1e-20), _clamp_min(s3 + s1, 1e-20)
    sin2_12, sin2_23, sin2_31 = ((s1 - s2) / d12) ** 2, ((s2 - s3) / d23) ** 2, ((s3 - s1) / d31) ** 2
    f_nfc = np.sqrt(sin2_12 + sin2_23 + sin2_31 + 1e-30) - k_nfc
    p_mean = _clamp_min((s1 + s2 + s3) / 3.0, 1e-20)
    F_shear = f_nfc * p_mean
    I1 = s[:, XX] + s[:, YY] + s[:, ZZ]
    F_tension = -I1 / 3.0 - sigma_t
    Fvp = np.maximum(F_shear, F_tension)
    is_tension = F_tension > F_shear
    dQ, _ = _dp_flow(s, I1, alpha_Q, is_tension)
    lam = _perzyna(Fvp, mu_1, N_1)
    return -dQ * lam[:, None], Fvp


class _StressOnly(_Element):
    """Viscoplastic elements whose rate depends on the stress only (no ISV)."""

    def tangent(self, sig, dt, theta, T):
        self.B = np.zeros((self.n, 6))
        self.G = fd_tangent(lambda s: self._rate(s)[0], sig)

    def eval_rate(self, sig, phi1, T):
        self.rate, self.Fvp = self._rate(sig)


class MohrCoulomb(_StressOnly):
    kind = "mohr_coulomb"

    def __init__(self, mu_1, N_1, cohesion, friction_angle, dilation_angle, sigma_t):
        a = lambda x: np.asarray(x, dtype=np.float64)
        super().__init__(len(mu_1))
        self.mu_1, self.N_1, self.sigma_t = a(mu_1), a(N_1), a(sigma_t)
        sin_phi, cos_phi, sin_psi = np.sin(a(friction_angle)), np.cos(a(friction_angle)), np.sin(a(dilation_angle))
        self.alpha_F = 2.0 * sin_phi / (np.sqrt(3.0) * (3.0 - sin_phi))            # :1644
        self.k_F = 6.0 * a(cohesion) * cos_phi / (np.sqrt(3.0) * (3.0 - sin_phi))  # :1645
        self.alpha_Q = 2.0 * sin_psi / (np.sqrt(3.0) * (3.0 - sin_psi))            # :1648
        self.Fvp = np.zeros(self.n)

    def _rate(self, sig):
        return rate_mohr_coulomb(sig, self.mu_1, self.N_1, self.alpha_F, self.k_F, self.alpha_Q, self.sigma_t)


class MatsuokaNakai(_StressOnly):
    kind = "matsuoka_nakai"

    def __init__(self, mu_1, N_1, cohesion, friction_angle, dilation_angle, sigma_t):
        a = lambda x: np.asarray(x, dtype=np.float64)
        super().__init__(len(mu_1))
        self.mu_1, self.N_1, self.sigma_t = a(mu_1), a(N_1), a(sigma_t)
        sin_phi, cos_phi, sin_psi = np.sin(a(friction_angle)), np.cos(a(friction_angle)), np.sin(a(dilation_angle))
        self.k_nfc = np.sqrt(2.0) * sin_phi                                        # :1816
        small = np.abs(sin_phi) < 1e-10
        safe = np.where(small, 1.0, sin_phi)
        self.shift = np.where(small, 0.0, a(cohesion) * cos_phi / safe)            # :1820-1826
        self.alpha_Q = 2.0 * sin_psi / (np.sqrt(3.0) * (3.0 - sin_psi))            # :1829
        self.Fvp = np.zeros(self.n)

    def _rate(self, sig):
        return rate_matsuoka_nakai(sig, self.mu_1, self.N_1, self.k_nfc, self.shift, self.alpha_Q, self.sigma_t)


class OracleMaterial:
    """Material (MaterialProps.py:22-331) + the constitutive glue of LinearMomentum
    (MomentumEquation.py:343-454, 799-890)."""

    def __init__(self, n):
        self.n = n
        self.C = np.zeros((n, 6, 6))
        self.C_inv = np.zeros((n, 6, 6))
        self.elems = []
        self.alpha_th = []      # list of (N,) arrays
        self.singular = 0

    def add_spring(self, E, nu):            # add_to_elastic, :125-148
        C = iso_matrix(E, nu)
        self.C = self.C + C
        self.C_inv = self.C_inv + np.linalg.inv(C)

    def add_thermoelastic(self, alpha):     # :161-170
        self.alpha_th.append(np.asarray(alpha, dtype=np.float64))

    def add(self, elem):                    # add_to_non_elastic, :150-159
        self.elems.append(elem)
        return elem

    # ---- tangent phase: LinearMomentum.compute_CT + compute_eps_rhs ------------------
    def tangent_phase(self, sig_k, T, T0, dt, theta):
        phi1, phi2 = dt * theta, dt * (1 - theta)
        G = np.zeros((self.n, 6, 6))
        B = np.zeros((self.n, 6))
        for e in self.elems:                # Material.compute_G_B, :172-200
            e.tangent(sig_k, dt, theta, T)
            G = G + e.G
            B = B + e.B
        self.G, self.B = G, B
        mat = self.C_inv + dt * (1 - theta) * G     # Material.compute_CT, :273-309
        try:
            self.CT = np.linalg.inv(mat)
        except np.linalg.LinAlgError:
            self.CT = np.linalg.inv(self.C_inv)
            self.singular = 0
            for i in range(self.n):
                try:
                    self.CT[i] = np.linalg.inv(mat[i])
                except np.linalg.LinAlgError:
                    self.singular += 1
        eps_ne_k = np.zeros((self.n, 6))    # MomentumEquation.py:359-377
        for e in self.elems:
            e.predictor(phi1, phi2)
            eps_ne_k = eps_ne_k + e.eps_k
        eps_th = np.zeros((self.n, 6))      # MomentumEquation.py:343-357
        dT = T - T0
        for a in self.alpha_th:
            eps_th[:, :3] = eps_th[:, :3] + (a * dT)[:, None] * 1.0
        # MomentumEquation.py:889
        self.eps_rhs = eps_ne_k + eps_th - dt * (1 - theta) * (B + ddot(G, sig_k))
        return self.CT, self.eps_rhs

    # ---- post-solve phase: compute_stress, increment ISVs, new rates -----------------
    def stress(self, eps):                  # MomentumEquation.py:844-866
        return ddot(self.CT, eps - self.eps_rhs)

    def elastic_stress(self, eps):          # MomentumEquation.py:822-842
        return ddot(self.C, eps)

    def post_phase(self, eps, sig_k, T, dt, theta):
        sig = self.stress(eps)
        for e in self.elems:                # Simulators.py:422
            e.increment_isv(sig, sig_k, dt)
        self.eval_rates(sig, dt * theta, T)  # Simulators.py:425
        return sig

    def eval_rates(self, sig, phi1, T):     # MomentumEquation.py:379-395
        for e in self.elems:
            e.eval_rate(sig, phi1, T)

    def commit(self, sig, sig_k, dt, theta):  # Simulators.py:509-517
        for e in self.elems:
            e.commit_isv()
        for e in self.elems:
            e.commit_rate()
        for e in self.elems:
            e.commit_strain(sig, sig_k, dt * (1 - theta))

    def commit_rates(self):                 # Simulators.py:365
        for e in self.elems:
            e.commit_rate()

    def snapshot(self):                     # MomentumEquation.py:456-477
        return [e.snapshot() for e in self.elems]

    def restore(self, snaps):               # MomentumEquation.py:479-494
        for e, s in zip(self.elems, snaps):
            e.restore(s)


def newton_error(eps_k, eps):
    """Simulators.py:433-436: ||eps_k - eps||_2 / ||eps||_2 over all 9 entries of all
    cells (off-diagonals counted twice), single rank (T9)."""
    w = np.array([1, 1, 1, 2, 2, 2], dtype=np.float64)
    num = np.sqrt((((eps_k - eps) ** 2) * w).sum())
    den = np.sqrt(((eps ** 2) * w).sum())
    with np.errstate(all="ignore"):
        return num / den
