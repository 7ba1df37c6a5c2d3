// constitutive.cuh — per-cell constitutive laws of SafeInCave, one thread per cell.
//
// Restates (from scratch, for the GPU) the element laws of the reference's
// safeincave/MaterialProps.py; every routine cites the lines it follows.  The arithmetic ORDER of
// every expression mirrors the reference's Python (left to right, one rounding per operation):
// this translation unit is compiled with -fmad=false and uses the bit-reproducible exp/pow of
// sic_math.h, so that the finite-difference tangents (amplification ~5e8) agree with the CPU
// oracle bit for bit.
#ifndef SIC_CONSTITUTIVE_CUH_
#define SIC_CONSTITUTIVE_CUH_

#include "sic_math.h"

#define SIC_R_GAS 8.32                    /* MaterialProps.py:915, 989 (sic: not 8.314) */
#define SIC_MPA 1.0e6                     /* Utils.py:35 */
#define SIC_SQRT27 5.196152422706632      /* np.sqrt(27) == 27**0.5 */
#define SIC_FD_EPS 1.0e-2                 /* MaterialProps.py:661 */

// a / C for a compile-time constant C, CORRECTLY ROUNDED without the ~15-instruction division sequence: with y = RN(1/C)
// (folded by the compiler), q = RN(a y) is a faithful quotient, r = a - C q is exact in one fma, and RN(q + r y) = RN(a / C)
// (Markstein's theorem; holds for every C whose significand is not all ones and every finite a with a normal quotient).
// Bit-identical to the oracle's true divisions; k_tangent spends 48 of these per cell and creep element.
#define SIC_DIV_CONST(a, C) sic_div_const((a), (C), 1.0 / (C))
__host__ __device__ __forceinline__ double sic_div_const(double a, double c, double rc) {
  const double q = a * rc;
  const double r = fma(-c, q, a);
  return fma(r, rc, q);
}
#define SIC_DESAI_EPS_STRESS 1.0e-1       /* MaterialProps.py:1460 */

namespace sic {

// sigma_v = C . eps_v, six products accumulated left to right (Utils.py:251-283)
__device__ __forceinline__ void ddot66(const double* __restrict__ C, const double e[6], double out[6]) {
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double r = C[i * 6 + 0] * e[0];
#pragma unroll
    for (int k = 1; k < 6; ++k) r = r + C[i * 6 + k] * e[k];
    out[i] = r;
  }
}

// isotropic matrix (c11 on the normal diagonal, c12 off-diagonal, c44 on the shear diagonal)
// times a Voigt vector: the zero products of the full 6x6 row leave the sum unchanged.
__device__ __forceinline__ void ddot_iso(double c11, double c12, double c44, const double e[6], double out[6]) {
  out[0] = (c11 * e[0] + c12 * e[1]) + c12 * e[2];
  out[1] = (c12 * e[0] + c11 * e[1]) + c12 * e[2];
  out[2] = (c12 * e[0] + c12 * e[1]) + c11 * e[2];
  out[3] = c44 * e[3];
  out[4] = c44 * e[4];
  out[5] = c44 * e[5];
}

// ---- DislocationCreep.compute_eps_ne_rate, MaterialProps.py:921-961 ------------------------
struct DislocationP { double A, Q, n; };
__device__ __forceinline__ void rate_dislocation(const double s[6], double T, const DislocationP& p, double out[6]) {
  double mean = SIC_DIV_CONST((s[0] + s[1]) + s[2], 3.0);
  double a = s[0] - s[1], b = s[0] - s[2], c = s[1] - s[2];
  double q = sqrt(0.5 * (((a * a + b * b) + c * c) + 6.0 * ((s[3] * s[3] + s[4] * s[4]) + s[5] * s[5])));
  double A_bar = (p.A * sic_exp(((-p.Q) / SIC_R_GAS) / T)) * sic_pow(q, p.n - 1.0);
  out[0] = A_bar * (s[0] - mean);
  out[1] = A_bar * (s[1] - mean);
  out[2] = A_bar * (s[2] - mean);
  out[3] = A_bar * s[3];
  out[4] = A_bar * s[4];
  out[5] = A_bar * s[5];
}

// ---- PressureSolutionCreep.compute_eps_ne_rate, MaterialProps.py:995-1034 ------------------
struct PressureSolP { double A, d, Q; };
__device__ __forceinline__ void rate_pressure_solution(const double s[6], double T, const PressureSolP& p, double out[6]) {
  double mean = SIC_DIV_CONST((s[0] + s[1]) + s[2], 3.0);
  double A_bar = ((p.A / ((p.d * p.d) * p.d)) / T) * sic_exp(((-p.Q) / SIC_R_GAS) / T);
  out[0] = A_bar * (s[0] - mean);
  out[1] = A_bar * (s[1] - mean);
  out[2] = A_bar * (s[2] - mean);
  out[3] = A_bar * s[3];
  out[4] = A_bar * s[4];
  out[5] = A_bar * s[5];
}

// ---- Viscoelastic (Kelvin-Voigt), MaterialProps.py:795-885 ---------------------------------
struct KelvinP { double eta, c11, c12, c44; };
// E = (eta I + phi2 C1)^-1 (:882-885) in closed form for the isotropic C1: returns its three
// distinct entries (normal diagonal, normal off-diagonal, shear diagonal).
__device__ __forceinline__ void kelvin_G(const KelvinP& p, double phi2, double& g11, double& g12, double& g44) {
  double a = p.eta + phi2 * p.c11;
  double b = phi2 * p.c12;
  double den = (a - b) * (a + 2.0 * b);
  g11 = (a + b) / den;
  g12 = -b / den;
  g44 = 1.0 / (p.eta + phi2 * p.c44);
}
// rate = G:(sigma - C1:(eps_old + phi1 rate_old))  (:855)
__device__ __forceinline__ void rate_kelvin(const double s[6], const KelvinP& p, double g11, double g12, double g44,
                                            const double eps_old[6], const double rate_old[6], double phi1,
                                            double out[6]) {
  double e[6], ce[6], d[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) e[k] = eps_old[k] + phi1 * rate_old[k];
  ddot_iso(p.c11, p.c12, p.c44, e, ce);
#pragma unroll
  for (int k = 0; k < 6; ++k) d[k] = s[k] - ce[k];
  ddot_iso(g11, g12, g44, d, out);
}

// ---- ViscoplasticDesai, MaterialProps.py:1037-1562 ------------------------------------------
struct DesaiP { double mu_1, N_1, a_1, eta, n, beta_1, beta, m, gamma, sigma_t; };

// compute_eps_ne_rate (:1291-1429) incl. extract_stress_components (:1199-1220),
// compute_stress_invariants (:1160-1197) and compute_Fvp (:1222-1246)
__device__ __forceinline__ void rate_desai(const double sig[6], double alpha, double alpha_0, const DesaiP& p,
                                           double rate[6], double& Fvp_out) {
  const double c13 = 1.0 / 3.0;
  double sxx = SIC_DIV_CONST(-sig[0], SIC_MPA), syy = SIC_DIV_CONST(-sig[1], SIC_MPA), szz = SIC_DIV_CONST(-sig[2], SIC_MPA);
  double sxy = SIC_DIV_CONST(-sig[3], SIC_MPA), sxz = SIC_DIV_CONST(-sig[4], SIC_MPA), syz = SIC_DIV_CONST(-sig[5], SIC_MPA);
  double I1 = (sxx + syy) + szz;
  double I2 = ((((sxx * syy + syy * szz) + sxx * szz) - sxy * sxy) - syz * syz) - sxz * sxz;
  double I3 = (((((sxx * syy) * szz + ((2.0 * sxy) * syz) * sxz) - szz * (sxy * sxy)) - sxx * (syz * syz)) -
               syy * (sxz * sxz));
  double J2 = c13 * (I1 * I1) - I2;
  double J3 = ((2.0 / 27.0) * ((I1 * I1) * I1) - (c13 * I1) * I2) + I3;
  const bool low_J2 = (J2 <= 1.0e-6);
  J2 = (J2 < 1.0e-6) ? 1.0e-6 : J2;  // clamp(min=): NaN stays NaN
  double J2_15 = sic_pow(J2, 1.5);
  double Sr = low_J2 ? 0.0 : (-(J3 * SIC_SQRT27)) / (2.0 * J2_15);
  double I1s = I1 + p.sigma_t;

  double powI1n = sic_pow(I1s, p.n);
  double I1s2 = I1s * I1s;
  double expb = sic_exp(p.beta_1 * I1s);
  double F2 = expb - p.beta * Sr;
  const bool low_F2 = (F2 < 1.0e-6);
  F2 = low_F2 ? 1.0e-6 : F2;
  double F2m = sic_pow(F2, p.m);
  double Fvp = J2 + (alpha * powI1n - p.gamma * I1s2) * F2m;
  Fvp_out = Fvp;

  double F1 = (-alpha) * powI1n + p.gamma * I1s2;
  double F2m1 = sic_pow(F2, p.m - 1.0);
  double dF1_dI1 = (2.0 * p.gamma) * I1s - (p.n * alpha) * sic_pow(I1s, p.n - 1.0);
  double dF2m_dI1 = ((p.beta_1 * p.m) * expb) * F2m1;
  double dF_dI1 = -(dF1_dI1 * F2m + F1 * dF2m_dI1);
  double dF2_dJ2 = (-(((3.0 * p.beta) * J3) * SIC_SQRT27)) / (4.0 * sic_pow(J2, 2.5));
  double dF_dJ2 = 1.0 - ((F1 * p.m) * F2m1) * dF2_dJ2;
  double dF_dJ3 = (((((-p.m) * F1) * p.beta) * SIC_SQRT27) * F2m1) / (2.0 * J2_15);

  double dI2[6] = {syy + szz, sxx + szz, sxx + syy, -2.0 * sxy, -2.0 * sxz, -2.0 * syz};
  double dI3[6] = {syy * szz - syz * syz, sxx * szz - sxz * sxz, sxx * syy - sxy * sxy,
                   2.0 * (sxz * syz - szz * sxy), 2.0 * (sxy * syz - syy * sxz), 2.0 * (sxz * sxy - sxx * syz)};
  double dJ2_dI1 = (2.0 / 3.0) * I1;
  double dJ3_dI1 = (2.0 / 9.0) * (I1 * I1) - c13 * I2;
  double dJ3_dI2 = (-c13) * I1;

  const bool off = low_J2 || low_F2 || (alpha <= 0.01 * alpha_0);
  double lam = 0.0;
  if (Fvp > 0.0) lam = p.mu_1 * sic_pow(Fvp / 1.0, p.N_1);
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double dI1k = (k < 3) ? 1.0 : 0.0;
    double dJ2_dS = dJ2_dI1 * dI1k + (-1.0) * dI2[k];
    double dJ3_dS = (dJ3_dI1 * dI1k + dJ3_dI2 * dI2[k]) + 1.0 * dI3[k];
    double dQ = (dF_dI1 * dI1k + dF_dJ2 * dJ2_dS) + dF_dJ3 * dJ3_dS;
    if (off) dQ = 0.0;
    rate[k] = (-dQ) * lam;
  }
}

// compute_residue (:1094-1117): returns r, writes qsi
__device__ __forceinline__ double desai_residue(const double rate[6], double alpha, double qsi_old, double alpha_0,
                                                double dt, const DesaiP& p, double& qsi) {
  double d = (rate[0] * rate[0] + rate[1] * rate[1]) + rate[2] * rate[2];
  double o = (rate[3] * rate[3] + rate[4] * rate[4]) + rate[5] * rate[5];
  qsi = qsi_old + sqrt(d + 2.0 * o) * dt;
  return alpha - p.a_1 / sic_pow(sic_pow(p.a_1 / alpha_0, 1.0 / p.eta) + qsi, p.eta);
}

// einsum('bij,bij->b') of two symmetric tensors in Voigt-6 (:1150)
__device__ __forceinline__ double ddot_sym(const double a[6], const double b[6]) {
  double d = (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
  double o = (a[3] * b[3] + a[4] * b[4]) + a[5] * b[5];
  return d + 2.0 * o;
}


// =============================================================================================
// MunsonDawsonCreep, MohrCoulombViscoplastic, MatsuokaNakaiViscoplastic
// =============================================================================================
#define SIC_MD_SQRT_EPS 1.4901161193847656e-8   /* MaterialProps.py:2013 */
#define SIC_MD_EPS_STRESS 1.0e-1               /* MaterialProps.py:2263 */
#define SIC_MD_H_MIN 1.0e-12                   /* MaterialProps.py:2283 */

// ---- MunsonDawsonCreep, MaterialProps.py:1971-2346 -------------------------------------------
struct MunsonDawsonP { double A, Q, n, K0, c, m, alpha_w, beta_w, delta, mu; };

__device__ __forceinline__ double clamp_min(double x, double lo) { return (x < lo) ? lo : x; }   // NaN stays NaN

// _compute_md_fields (:2104-2167): deviator, von Mises stress (floored at 1 Pa), steady-state rate,
// transient limit eps_t*, transient function F
__device__ __forceinline__ void md_fields(const double s[6], double T, double zeta, const MunsonDawsonP& p,
                                          double dev[6], double& sigma_safe, double& epsdot_ss, double& ets, double& F) {
  const double mean = SIC_DIV_CONST((s[0] + s[1]) + s[2], 3.0);
  dev[0] = s[0] - mean; dev[1] = s[1] - mean; dev[2] = s[2] - mean;
  dev[3] = s[3]; dev[4] = s[4]; dev[5] = s[5];
  const double a = s[0] - s[1], b = s[0] - s[2], c = s[1] - s[2];
  const double sigma = sqrt(0.5 * (((a * a + b * b) + c * c) + 6.0 * ((s[3] * s[3] + s[4] * s[4]) + s[5] * s[5])));
  sigma_safe = clamp_min(sigma, 1.0);
  const double mu_safe = clamp_min(p.mu, 1.0);
  epsdot_ss = (p.A * sic_exp((-p.Q) / (SIC_R_GAS * T))) * sic_pow(sigma_safe, p.n);
  const double ratio = clamp_min(sigma_safe / mu_safe, 1.0e-30);
  ets = (p.K0 * sic_exp(p.c * T)) * sic_pow(ratio, p.m);
  ets = clamp_min(ets, 1.0e-50);
  const double Delta = p.alpha_w + p.beta_w * sic_log10(ratio);
  const double r_arg = 1.0 - (zeta / ets);
  const double r2 = r_arg * r_arg;
  double arg = (zeta <= ets) ? Delta * r2 : (-p.delta) * r2;
  arg = (arg < -50.0) ? -50.0 : ((arg > 50.0) ? 50.0 : arg);
  F = sic_exp(arg);
}

// compute_eps_ne_rate (:2187-2229)
__device__ __forceinline__ void rate_munson_dawson(const double s[6], double T, double zeta, const MunsonDawsonP& p,
                                                   double rate[6], double& F, double& ets) {
  double dev[6], sigma_safe, epsdot_ss;
  md_fields(s, T, zeta, p, dev, sigma_safe, epsdot_ss, ets, F);
  const double scalar_rate = F * epsdot_ss;
  const double f = 1.5 / sigma_safe;
#pragma unroll
  for (int k = 0; k < 6; ++k) rate[k] = (f * dev[k]) * scalar_rate;
}

// compute_residue (:2169-2181)
__device__ __forceinline__ double md_residue(const double s[6], double T, double zeta, double zeta_old, double dt,
                                             const MunsonDawsonP& p) {
  double dev[6], sigma_safe, epsdot_ss, ets, F;
  md_fields(s, T, zeta, p, dev, sigma_safe, epsdot_ss, ets, F);
  return (zeta - zeta_old) - ((F - 1.0) * epsdot_ss) * dt;
}

// ---- Drucker-Prager flow direction shared by MohrCoulomb / MatsuokaNakai (:1705-1731, 1927-1953) ------
__device__ __forceinline__ void dp_flow_rate(const double c[6], double I1, double alpha_Q, bool is_tension, double lam,
                                             double rate[6]) {
  const double I2 = ((((c[0] * c[1] + c[1] * c[2]) + c[0] * c[2]) - c[3] * c[3]) - c[5] * c[5]) - c[4] * c[4];
  const double J2 = clamp_min((1.0 / 3.0) * (I1 * I1) - I2, 1.0e-20);
  const double inv = 1.0 / (2.0 * sqrt(J2));
  double dQ[6];
  dQ[0] = inv * ((2.0 / 3.0) * I1 - (c[1] + c[2])) - alpha_Q;
  dQ[1] = inv * ((2.0 / 3.0) * I1 - (c[0] + c[2])) - alpha_Q;
  dQ[2] = inv * ((2.0 / 3.0) * I1 - (c[0] + c[1])) - alpha_Q;
  dQ[3] = inv * (2.0 * c[3]);
  dQ[4] = inv * (2.0 * c[4]);
  dQ[5] = inv * (2.0 * c[5]);
  if (is_tension) {
    dQ[0] = dQ[1] = dQ[2] = -1.0 / 3.0;
    dQ[3] = dQ[4] = dQ[5] = 0.0;
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) rate[k] = (-dQ[k]) * lam;
}

// ---- MohrCoulombViscoplastic.compute_eps_ne_rate, MaterialProps.py:1652-1746 ------------------
struct MohrCoulombP { double mu_1, N_1, alpha_F, k_F, alpha_Q, sigma_t; };
__device__ __forceinline__ void rate_mohr_coulomb(const double sig[6], const MohrCoulombP& p, double rate[6],
                                                  double& Fvp_out) {
  double c[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) c[k] = SIC_DIV_CONST(-sig[k], SIC_MPA);
  const double I1 = (c[0] + c[1]) + c[2];
  const double I2 = ((((c[0] * c[1] + c[1] * c[2]) + c[0] * c[2]) - c[3] * c[3]) - c[5] * c[5]) - c[4] * c[4];
  const double J2 = clamp_min((1.0 / 3.0) * (I1 * I1) - I2, 1.0e-20);
  const double F_shear = (sqrt(J2) - p.alpha_F * I1) - p.k_F;
  const double F_tension = SIC_DIV_CONST(-I1, 3.0) - p.sigma_t;
  const double Fvp = (F_shear > F_tension || F_shear != F_shear) ? F_shear : F_tension;   // torch.maximum: NaN wins
  Fvp_out = (F_tension != F_tension) ? F_tension : Fvp;
  double lam = 0.0;
  if (Fvp_out > 0.0) lam = p.mu_1 * sic_pow(Fvp_out / 1.0, p.N_1);
  dp_flow_rate(c, I1, p.alpha_Q, F_tension > F_shear, lam, rate);
}

// ---- eigenvalues of a symmetric 3x3 (ascending): six cyclic Jacobi sweeps, fixed operation sequence ----
// (torch.linalg.eigvalsh in the reference, MaterialProps.py:1882; the oracle restates THIS routine
// operation for operation next to its LAPACK call so that device and oracle agree bit for bit)
__device__ __forceinline__ void jacobi_rotate(double& app, double& aqq, double& apq, double& arp, double& arq) {
  double t = 0.0;
  if (apq != 0.0) {
    const double theta = (aqq - app) / (2.0 * apq);
    const double at = fabs(theta);
    t = 1.0 / (at + sqrt(theta * theta + 1.0));
    if (theta < 0.0) t = -t;
  }
  const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
  app = app - t * apq;
  aqq = aqq + t * apq;
  apq = 0.0;
  const double rp = c * arp - s * arq, rq = s * arp + c * arq;
  arp = rp; arq = rq;
}
__device__ __forceinline__ void eigvals_sym3(const double c[6], double& e_lo, double& e_mid, double& e_hi) {
  double a00 = c[0], a11 = c[1], a22 = c[2], a01 = c[3], a02 = c[4], a12 = c[5];
#pragma unroll 1
  for (int sweep = 0; sweep < 6; ++sweep) {
    jacobi_rotate(a00, a11, a01, a02, a12);   // (p,q) = (0,1), r = 2
    jacobi_rotate(a00, a22, a02, a01, a12);   // (0,2), r = 1
    jacobi_rotate(a11, a22, a12, a01, a02);   // (1,2), r = 0
  }
  double x = a00, y = a11, z = a22, t;
  if (y < x) { t = x; x = y; y = t; }
  if (z < y) { t = y; y = z; z = t; }
  if (y < x) { t = x; x = y; y = t; }
  e_lo = x; e_mid = y; e_hi = z;
}

// ---- MatsuokaNakaiViscoplastic.compute_eps_ne_rate, MaterialProps.py:1835-1968 ------------------
struct MatsuokaNakaiP { double mu_1, N_1, k_nfc, shift, alpha_Q, sigma_t; };
__device__ __forceinline__ void rate_matsuoka_nakai(const double sig[6], const MatsuokaNakaiP& p, double rate[6],
                                                    double& Fvp_out) {
  double c[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) c[k] = SIC_DIV_CONST(-sig[k], SIC_MPA);
  double sig3, sig2, sig1;
  eigvals_sym3(c, sig3, sig2, sig1);
  const double s1 = sig1 + p.shift, s2 = sig2 + p.shift, s3 = sig3 + p.shift;
  const double d12 = clamp_min(s1 + s2, 1.0e-20), d23 = clamp_min(s2 + s3, 1.0e-20), d31 = clamp_min(s3 + s1, 1.0e-20);
  const double q12 = (s1 - s2) / d12, q23 = (s2 - s3) / d23, q31 = (s3 - s1) / d31;
  const double f_nfc = sqrt(((q12 * q12 + q23 * q23) + q31 * q31) + 1.0e-30) - p.k_nfc;
  const double p_mean = clamp_min(SIC_DIV_CONST((s1 + s2) + s3, 3.0), 1.0e-20);
  const double F_shear = f_nfc * p_mean;
  const double I1 = (c[0] + c[1]) + c[2];
  const double F_tension = SIC_DIV_CONST(-I1, 3.0) - p.sigma_t;
  const double Fvp = (F_shear > F_tension || F_shear != F_shear) ? F_shear : F_tension;
  Fvp_out = (F_tension != F_tension) ? F_tension : Fvp;
  double lam = 0.0;
  if (Fvp_out > 0.0) lam = p.mu_1 * sic_pow(Fvp_out / 1.0, p.N_1);
  dp_flow_rate(c, I1, p.alpha_Q, F_tension > F_shear, lam, rate);
}

// ---- NonElasticElement.compute_E, MaterialProps.py:640-675 ---------------------------------
// Central FD with an ABSOLUTE step of 1e-2 Pa on a running copy (+=, -=, -=, +=),
// shear columns doubled (T3).  consume(k, col) receives column k of E.  The k loop is kept
// rolled (two inlined rate evaluations per trip); the perturbed component is selected by
// predication so that s[] keeps static register indices.
template <bool UNROLL = false, class RateFn, class ColFn>
__device__ __forceinline__ void fd_columns(RateFn rate_fn, const double sig[6], ColFn consume) {
  double s[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) s[c] = sig[c];
  auto column = [&](int k) {
    double ra[6], rb[6], col[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) s[c] = (c == k) ? s[c] + SIC_FD_EPS : s[c];
    rate_fn(s, ra);
#pragma unroll
    for (int c = 0; c < 6; ++c) s[c] = (c == k) ? (s[c] - SIC_FD_EPS) - SIC_FD_EPS : s[c];
    rate_fn(s, rb);
#pragma unroll
    for (int c = 0; c < 6; ++c) s[c] = (c == k) ? s[c] + SIC_FD_EPS : s[c];
    const double phi = (k < 3) ? 1.0 : 2.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) col[i] = SIC_DIV_CONST(phi * (ra[i] - rb[i]), 2.0 * SIC_FD_EPS);
    consume(k, col);
  };
  // UNROLL (cheap rate laws in the creep-only kernels): six inlined copies, k is a compile-time constant in each, so
  // the selects that keep s[] and the consumer's 6x6 in registers under a runtime k (108 per column) disappear
  if constexpr (UNROLL) {
#pragma unroll
    for (int k = 0; k < 6; ++k) column(k);
  } else {
#pragma unroll 1
    for (int k = 0; k < 6; ++k) column(k);
  }
}

// In-place 6x6 inverse, Gauss-Jordan with partial pivoting, static register indexing.
// Returns false when a pivot is exactly zero (torch.linalg.inv raises LinAlgError there,
// MaterialProps.py:293-309 -> elastic fallback).
__device__ __forceinline__ bool inverse6(double a[36]) {
  int piv[6];
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    int p = k;
    double best = fabs(a[k * 6 + k]);
#pragma unroll
    for (int j = k + 1; j < 6; ++j) {
      double v = fabs(a[j * 6 + k]);
      if (v > best) { best = v; p = j; }
    }
    piv[k] = p;
    if (best == 0.0) ok = false;
#pragma unroll
    for (int j = k + 1; j < 6; ++j) {
      if (p == j) {
#pragma unroll
        for (int c = 0; c < 6; ++c) { double t = a[k * 6 + c]; a[k * 6 + c] = a[j * 6 + c]; a[j * 6 + c] = t; }
      }
    }
    double d = 1.0 / a[k * 6 + k];
    a[k * 6 + k] = 1.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) a[k * 6 + c] *= d;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (i != k) {
        double f = a[i * 6 + k];
        a[i * 6 + k] = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) a[i * 6 + c] = fma(-f, a[k * 6 + c], a[i * 6 + c]);
      }
    }
  }
#pragma unroll
  for (int k = 5; k >= 0; --k) {
#pragma unroll
    for (int j = k + 1; j < 6; ++j) {
      if (piv[k] == j) {
#pragma unroll
        for (int r = 0; r < 6; ++r) { double t = a[r * 6 + k]; a[r * 6 + k] = a[r * 6 + j]; a[r * 6 + j] = t; }
      }
    }
  }
  return ok;
}

}  // namespace sic
#endif  // SIC_CONSTITUTIVE_CUH_
