// constitutive.cu — hot-path part (1): per-cell constitutive update kernels (FP64, sm_100a).
// One thread per cell; SoA Voigt components so that every warp load/store is a full coalesced
// 256-byte row segment.  Compiled with -fmad=false (see constitutive.cuh).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/safeincave_cuda.h"
#include "common.cuh"
#include "constitutive.cuh"

namespace sic {

#define SIC_CELL_THREADS 128

struct Row {
  const double* __restrict__ p;
  __device__ __forceinline__ double operator[](int k) const { return __ldg(p + k); }
};

__device__ __forceinline__ void load6(const double* __restrict__ a, int ns, int i, double v[6]) {
#pragma unroll
  for (int c = 0; c < 6; ++c) v[c] = a[(size_t)c * ns + i];
}
__device__ __forceinline__ void store6(double* __restrict__ a, int ns, int i, const double v[6]) {
#pragma unroll
  for (int c = 0; c < 6; ++c) a[(size_t)c * ns + i] = v[c];
}

__device__ __forceinline__ DesaiP load_desai(const Row& row, int off) {
  DesaiP q;
  q.mu_1 = row[off + 0]; q.N_1 = row[off + 1]; q.a_1 = row[off + 2]; q.eta = row[off + 3];
  q.n = row[off + 4]; q.beta_1 = row[off + 5]; q.beta = row[off + 6]; q.m = row[off + 7];
  q.gamma = row[off + 8]; q.sigma_t = row[off + 9];
  return q;
}

__device__ __forceinline__ MunsonDawsonP load_md(const Row& row, int off) {
  MunsonDawsonP q;
  q.A = row[off + 0]; q.Q = row[off + 1]; q.n = row[off + 2]; q.K0 = row[off + 3]; q.c = row[off + 4];
  q.m = row[off + 5]; q.alpha_w = row[off + 6]; q.beta_w = row[off + 7]; q.delta = row[off + 8]; q.mu = row[off + 9];
  return q;
}
__device__ __forceinline__ MohrCoulombP load_mc(const Row& row, int off) {
  return MohrCoulombP{row[off], row[off + 1], row[off + 2], row[off + 3], row[off + 4], row[off + 5]};
}
__device__ __forceinline__ MatsuokaNakaiP load_mn(const Row& row, int off) {
  return MatsuokaNakaiP{row[off], row[off + 1], row[off + 2], row[off + 3], row[off + 4], row[off + 5]};
}

// Build variants of k_tangent (scripts/build_variants.py times them against each other on the GPU):
//   SIC_TAN_SMEM_G    1: the accumulated 6x6 G lives in shared memory while the elements' finite-difference columns are
//                        formed (dynamic column index for free, 72 registers fewer during the rate evaluations); it is
//                        pulled into registers only for the inverse.  0: in registers, columns added by predication.
//   SIC_TAN_UNROLL    1: the six FD columns of DislocationCreep / PressureSolutionCreep as six inlined copies (SET 0).
//   SIC_TAN_MINBLOCKS resident CTAs per SM the SET-0 instantiation is compiled for (register cap 65536 / 128 / that).
#ifndef SIC_TAN_SMEM_G
#define SIC_TAN_SMEM_G 1
#endif
#ifndef SIC_TAN_UNROLL
#define SIC_TAN_UNROLL 0
#endif
#ifndef SIC_TAN_MINBLOCKS
#define SIC_TAN_MINBLOCKS 4
#endif

// the 6x6 in shared memory, one column of the array per thread: entry j of this thread at base[j * SIC_CELL_THREADS]
struct SmemMat {
  double* base;
  __device__ __forceinline__ double& operator[](int j) const { return base[j * SIC_CELL_THREADS]; }
};
__device__ __forceinline__ void add_col(const SmemMat& G, int k, const double col[6]) {
#pragma unroll
  for (int i = 0; i < 6; ++i) G[i * 6 + k] = G[i * 6 + k] + col[i];
}

// accumulate column k of a 6x6 held in registers without dynamic indexing
__device__ __forceinline__ void add_col(double G[36], int k, const double col[6]) {
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int c = 0; c < 6; ++c) G[i * 6 + c] = (c == k) ? G[i * 6 + c] + col[i] : G[i * 6 + c];
  }
}

// Desai: compute_B_and_H_over_h (MaterialProps.py:1432-1500).  Leaves r, h, h_small, P, Q, qsi.
struct DesaiLin { double r, h, qsi, P[6], Q[6]; bool h_small; };

__device__ __forceinline__ void desai_linearise(const double sig_k[6], const double rate_cur[6], double alpha,
                                                double alpha_0, double qsi_old, double dt, const DesaiP& dp,
                                                DesaiLin& L) {
  double eps_alpha = 0.0001 * alpha;
  double alpha_eps = alpha + eps_alpha;
  double rate_eps[6], fv;
  rate_desai(sig_k, alpha_eps, alpha_0, dp, rate_eps, fv);
  double qsi;
  L.r = desai_residue(rate_cur, alpha, qsi_old, alpha_0, dt, dp, qsi);
  double r_eps = desai_residue(rate_eps, alpha_eps, qsi_old, alpha_0, dt, dp, qsi);
  double h = (r_eps - L.r) / eps_alpha;
#pragma unroll
  for (int c = 0; c < 6; ++c) L.Q[c] = (rate_eps[c] - rate_cur[c]) / eps_alpha;
  L.h_small = fabs(h) < 1.0e-6;
  if (L.h_small) h = 1.0;
  L.h = h;
  double s[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) s[c] = sig_k[c];
#pragma unroll 1
  for (int k = 0; k < 6; ++k) {
#pragma unroll
    for (int c = 0; c < 6; ++c) s[c] = (c == k) ? s[c] + SIC_DESAI_EPS_STRESS : s[c];
    double rp[6];
    rate_desai(s, alpha, alpha_0, dp, rp, fv);
    double r_p = desai_residue(rp, alpha, qsi_old, alpha_0, dt, dp, qsi);
    double Pk = SIC_DIV_CONST(r_p - L.r, SIC_DESAI_EPS_STRESS);
#pragma unroll
    for (int c = 0; c < 6; ++c) L.P[c] = (c == k) ? Pk : L.P[c];
#pragma unroll
    for (int c = 0; c < 6; ++c) s[c] = (c == k) ? s[c] - SIC_DESAI_EPS_STRESS : s[c];
  }
  L.qsi = qsi;  // whatever the LAST residue call left behind
  if (L.h_small) {
#pragma unroll
    for (int c = 0; c < 6; ++c) L.P[c] = 0.0;
  }
}

// column k of H/h (compute_H :1503-1562: H = Q (x) P with the shear COLUMNS doubled)
__device__ __forceinline__ double desai_Hh(const DesaiLin& L, int i, int k) {
  if (L.h_small) return 0.0;
  double Pk = 0.0;
#pragma unroll
  for (int c = 0; c < 6; ++c) Pk = (c == k) ? L.P[c] : Pk;
  double H = (k < 3) ? L.Q[i] * Pk : (2.0 * L.Q[i]) * Pk;
  return H / L.h;
}

// Munson-Dawson: compute_B_and_H_over_h (MaterialProps.py:2235-2313).  Leaves r, h, h_small, P, Q in the same
// record the Desai element uses (H/h has the same rank-one structure, desai_Hh).
__device__ __forceinline__ void md_linearise(const double sig_k[6], double T, double zeta, double zeta_old, double dt,
                                             const MunsonDawsonP& mp, DesaiLin& L) {
  double rate_ref[6], rate_z[6], Fd, ets_now;
  rate_munson_dawson(sig_k, T, zeta, mp, rate_ref, Fd, ets_now);
  const double zeta_scale = clamp_min(fabs(zeta) + ets_now, 1.0e-30);
  const double eps_zeta = SIC_MD_SQRT_EPS * zeta_scale;
  L.r = md_residue(sig_k, T, zeta, zeta_old, dt, mp);
  const double zeta_eps = zeta + eps_zeta;
  const double r_zeta = md_residue(sig_k, T, zeta_eps, zeta_old, dt, mp);
  double h = (r_zeta - L.r) / eps_zeta;
  double ets_z;
  rate_munson_dawson(sig_k, T, zeta_eps, mp, rate_z, Fd, ets_z);
#pragma unroll
  for (int c = 0; c < 6; ++c) L.Q[c] = (rate_z[c] - rate_ref[c]) / eps_zeta;
  L.h_small = fabs(h) < SIC_MD_H_MIN;
  if (L.h_small) h = 1.0;
  L.h = h;
  double s[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) s[c] = sig_k[c];
#pragma unroll 1
  for (int k = 0; k < 6; ++k) {
#pragma unroll
    for (int c = 0; c < 6; ++c) s[c] = (c == k) ? s[c] + SIC_MD_EPS_STRESS : s[c];
    const double r_sig = md_residue(s, T, zeta, zeta_old, dt, mp);
    const double Pk = SIC_DIV_CONST(r_sig - L.r, SIC_MD_EPS_STRESS);
#pragma unroll
    for (int c = 0; c < 6; ++c) L.P[c] = (c == k) ? Pk : L.P[c];
#pragma unroll
    for (int c = 0; c < 6; ++c) s[c] = (c == k) ? s[c] - SIC_MD_EPS_STRESS : s[c];
  }
  L.qsi = 0.0;
  if (L.h_small) {
#pragma unroll
    for (int c = 0; c < 6; ++c) L.P[c] = 0.0;
  }
}

// =============================================================================================
// tangent phase: LinearMomentum.compute_CT + compute_eps_rhs (MomentumEquation.py:799-820, 868-890)
// SET (sic_element_set): 0 = Kelvin / DislocationCreep / PressureSolutionCreep, 1 = + ViscoplasticDesai (together the
// elements of BASELINE configs 1-4), 2 = + the SURVEY 8f elements; the host picks the instantiation from the material.
// =============================================================================================
template <int SET>
__global__ void __launch_bounds__(SIC_CELL_THREADS, SET == 0 ? SIC_TAN_MINBLOCKS : 1) k_tangent(sic_problem_t P, double dt, double theta) {
  constexpr bool DESAI = SET >= 1, EXT = SET >= 2;   // element set of the instantiation (see sic_element_set)
#if SIC_TAN_SMEM_G
  __shared__ double g_sm[36 * SIC_CELL_THREADS];       // 36 KB: no barrier anywhere, every thread owns its column
#endif
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_cells) return;
  const int ns = P.cell_stride;
  const double phi1 = dt * theta;
  const double phi2 = dt * (1.0 - theta);
  Row row{P.mat_table + (size_t)P.mat_id[i] * P.row_len};
  const double T = P.T[i];

  double sk[6];
  load6(P.sig_k, ns, i, sk);

#if SIC_TAN_SMEM_G
  const SmemMat G{g_sm + threadIdx.x};
#else
  double G[36];
#endif
#pragma unroll
  for (int j = 0; j < 36; ++j) G[j] = 0.0;
  double B[6] = {0, 0, 0, 0, 0, 0};
  double eps_ne_k[6] = {0, 0, 0, 0, 0, 0};

  for (int e = 0; e < P.n_elems; ++e) {
    const sic_elem_t& el = P.elems[e];
    const int off = el.param_off;
    double eps_old[6], rate_old[6], rate_cur[6];
    load6(el.eps_old, ns, i, eps_old);
    load6(el.rate_old, ns, i, rate_old);
    load6(el.rate, ns, i, rate_cur);
    // NonElasticElement.compute_eps_ne_k (MaterialProps.py:586-605)
    double ek[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) ek[c] = (eps_old[c] + phi1 * rate_old[c]) + phi2 * rate_cur[c];
    store6(el.eps_k, ns, i, ek);
#pragma unroll
    for (int c = 0; c < 6; ++c) eps_ne_k[c] = eps_ne_k[c] + ek[c];

    if (el.kind == SIC_ELEM_KELVIN) {
      KelvinP kp{row[off], row[off + 1], row[off + 2], row[off + 3]};
      double g11, g12, g44;
      kelvin_G(kp, phi2, g11, g12, g44);
      G[0] += g11; G[7] += g11; G[14] += g11;
      G[1] += g12; G[2] += g12; G[6] += g12; G[8] += g12; G[12] += g12; G[13] += g12;
      G[21] += g44; G[28] += g44; G[35] += g44;
    } else if (el.kind == SIC_ELEM_DISLOCATION) {
      DislocationP dp{row[off], row[off + 1], row[off + 2]};
      fd_columns<SET == 0 && SIC_TAN_UNROLL>([&](const double* s, double* r) { rate_dislocation(s, T, dp, r); }, sk,
                           [&](int k, const double* col) { add_col(G, k, col); });
    } else if (el.kind == SIC_ELEM_PRESSURE_SOL) {
      PressureSolP pp{row[off], row[off + 1], row[off + 2]};
      fd_columns<SET == 0 && SIC_TAN_UNROLL>([&](const double* s, double* r) { rate_pressure_solution(s, T, pp, r); }, sk,
                           [&](int k, const double* col) { add_col(G, k, col); });
    } else if (DESAI && el.kind == SIC_ELEM_DESAI) {
      DesaiP dp = load_desai(row, off);
      double* ds = el.desai;
      const double alpha = ds[(size_t)SIC_DS_ALPHA * ns + i];
      const double alpha_0 = ds[(size_t)SIC_DS_ALPHA0 * ns + i];
      const double qsi_old = ds[(size_t)SIC_DS_QSI_OLD * ns + i];
      DesaiLin L;
      desai_linearise(sk, rate_cur, alpha, alpha_0, qsi_old, dt, dp, L);
      ds[(size_t)SIC_DS_QSI * ns + i] = L.qsi;
      ds[(size_t)SIC_DS_R * ns + i] = L.r;
      ds[(size_t)SIC_DS_H * ns + i] = L.h;
      ds[(size_t)SIC_DS_HSMALL * ns + i] = L.h_small ? 1.0 : 0.0;
      ds[(size_t)SIC_DS_ALPHA_K * ns + i] = alpha;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        ds[(size_t)(SIC_DS_P + c) * ns + i] = L.P[c];
        ds[(size_t)(SIC_DS_Q + c) * ns + i] = L.Q[c];
      }
      if (!L.h_small) {
        double rh = L.r / L.h;
#pragma unroll
        for (int c = 0; c < 6; ++c) B[c] = B[c] + rh * L.Q[c];
      }
      fd_columns([&](const double* s, double* r) { double fv; rate_desai(s, alpha, alpha_0, dp, r, fv); }, sk,
                 [&](int k, const double* col) {
                   double gc[6];
#pragma unroll
                   for (int r = 0; r < 6; ++r) gc[r] = col[r] - desai_Hh(L, r, k);
                   add_col(G, k, gc);
                 });
    }
    if constexpr (EXT) {
      if (el.kind == SIC_ELEM_MUNSON_DAWSON) {
        MunsonDawsonP mp = load_md(row, off);
        double* ms = el.desai;
        const double zeta = ms[(size_t)SIC_MD_ZETA * ns + i];
        const double zeta_old = ms[(size_t)SIC_MD_ZETA_OLD * ns + i];
        DesaiLin L;
        md_linearise(sk, T, zeta, zeta_old, dt, mp, L);
        ms[(size_t)SIC_MD_R * ns + i] = L.r;
        ms[(size_t)SIC_MD_H * ns + i] = L.h;
        ms[(size_t)SIC_MD_HSMALL * ns + i] = L.h_small ? 1.0 : 0.0;
        ms[(size_t)SIC_MD_ZETA_K * ns + i] = zeta;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          ms[(size_t)(SIC_MD_P + c) * ns + i] = L.P[c];
          ms[(size_t)(SIC_MD_Q + c) * ns + i] = L.Q[c];
        }
        if (!L.h_small) {
          double rh = L.r / L.h;
#pragma unroll
          for (int c = 0; c < 6; ++c) B[c] = B[c] + rh * L.Q[c];
        }
        fd_columns([&](const double* s, double* r) { double f, e; rate_munson_dawson(s, T, zeta, mp, r, f, e); }, sk,
                   [&](int k, const double* col) {
                     double gc[6];
#pragma unroll
                     for (int r = 0; r < 6; ++r) gc[r] = col[r] - desai_Hh(L, r, k);
                     add_col(G, k, gc);
                   });
      } else if (el.kind == SIC_ELEM_MOHR_COULOMB) {
        MohrCoulombP cp = load_mc(row, off);
        fd_columns([&](const double* s, double* r) { double f; rate_mohr_coulomb(s, cp, r, f); }, sk,
                   [&](int k, const double* col) { add_col(G, k, col); });
      } else if (el.kind == SIC_ELEM_MATSUOKA_NAKAI) {
        MatsuokaNakaiP np_ = load_mn(row, off);
        fd_columns([&](const double* s, double* r) { double f; rate_matsuoka_nakai(s, np_, r, f); }, sk,
                   [&](int k, const double* col) { add_col(G, k, col); });
      }
    }
  }

  // eps_th = sum alpha_th (T - T0) I   (MaterialProps.py:365-382, MomentumEquation.py:343-357)
  double eps_th = 0.0;
  if (P.n_thermo > 0) {
    const double dT = T - P.T0[i];
    for (int a = 0; a < P.n_thermo; ++a) eps_th = eps_th + (row[P.thermo_off + a] * dT) * 1.0;
  }

  // eps_rhs = eps_ne_k + eps_th - phi2 (B + G:sigma_k)   (MomentumEquation.py:889)
  double A[36];                 // G in registers from here on (the inverse indexes it statically)
#pragma unroll
  for (int j = 0; j < 36; ++j) A[j] = G[j];
  double Gs[6], er[6];
  ddot66(A, sk, Gs);
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    double th = (c < 3) ? eps_th : 0.0;
    er[c] = (eps_ne_k[c] + th) - phi2 * (B[c] + Gs[c]);
  }
  store6(P.eps_rhs, ns, i, er);

  // CT = inv(C_inv + phi2 G)   (Material.compute_CT, MaterialProps.py:273-309)
  const int so = P.spring_off;
  const double ci11 = row[so + 3], ci12 = row[so + 4], ci44 = row[so + 5];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double ci = 0.0;
      if (r == c) ci = (r < 3) ? ci11 : ci44;
      else if (r < 3 && c < 3) ci = ci12;
      A[r * 6 + c] = ci + phi2 * A[r * 6 + c];
    }
  }
  bool ok = inverse6(A);
  if (!ok) {  // singular -> elastic tangent for this cell (MaterialProps.py:296-309)
    const double c11 = row[so], c12 = row[so + 1], c44 = row[so + 2];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double v = 0.0;
        if (r == c) v = (r < 3) ? c11 : c44;
        else if (r < 3 && c < 3) v = c12;
        A[r * 6 + c] = v;
      }
    }
    if (P.n_singular) atomicAdd(P.n_singular, 1);
  }
#pragma unroll
  for (int j = 0; j < 36; ++j) P.CT[SIC_CT_INDEX(j, i)] = A[j];
}

// CT <- C, eps_rhs <- 0 (operator of solve_elastic_response, MomentumEquation.py:892-923)
__global__ void __launch_bounds__(SIC_CELL_THREADS) k_elastic_tangent(sic_problem_t P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_cells) return;
  const int ns = P.cell_stride;
  Row row{P.mat_table + (size_t)P.mat_id[i] * P.row_len};
  const int so = P.spring_off;
  const double c11 = row[so], c12 = row[so + 1], c44 = row[so + 2];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double v = 0.0;
      if (r == c) v = (r < 3) ? c11 : c44;
      else if (r < 3 && c < 3) v = c12;
      P.CT[SIC_CT_INDEX(r * 6 + c, i)] = v;
    }
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) P.eps_rhs[(size_t)c * ns + i] = 0.0;
}

// =============================================================================================
// post-solve phase of one Newton iteration (Simulators.py:416-436)
// =============================================================================================
template <int SET>
__global__ void __launch_bounds__(SIC_CELL_THREADS) k_post(sic_problem_t P, const double* __restrict__ u, double dt,
                                                          double theta, double kelvin_phi2, int flags,
                                                          double* __restrict__ err_scratch) {
  constexpr bool DESAI = SET >= 1, EXT = SET >= 2;   // element set of the instantiation (see sic_element_set)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int ns = P.cell_stride;
  const bool live = i < P.n_cells;
  double e_num = 0.0, e_den = 0.0;
  if (live) {
    double eps[6], sig[6];
    if (flags & SIC_POST_STRAIN) {
      // eps = sym(sum_a u_a (x) grad phi_a): exact DG0 interpolation of eps(u) for P1
      // (compute_total_strain, MomentumEquation.py:326-341; Utils.py:83-136)
      double g[12], ua[12];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int node = P.conn[(size_t)a * ns + i];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          g[3 * a + j] = P.grad[(size_t)(3 * a + j) * ns + i];
          ua[3 * a + j] = u[3 * (size_t)node + j];
        }
      }
      double exx = 0, eyy = 0, ezz = 0, exy = 0, exz = 0, eyz = 0;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const double gx = g[3 * a], gy = g[3 * a + 1], gz = g[3 * a + 2];
        const double ux = ua[3 * a], uy = ua[3 * a + 1], uz = ua[3 * a + 2];
        exx += ux * gx; eyy += uy * gy; ezz += uz * gz;
        exy += ux * gy + uy * gx; exz += ux * gz + uz * gx; eyz += uy * gz + uz * gy;
      }
      eps[0] = exx; eps[1] = eyy; eps[2] = ezz; eps[3] = 0.5 * exy; eps[4] = 0.5 * exz; eps[5] = 0.5 * eyz;
      if (flags & SIC_POST_ERROR) {
        double ep[6];
        load6(P.eps_prev, ns, i, ep);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const double w = (c < 3) ? 1.0 : 2.0;
          const double d = ep[c] - eps[c];
          e_num += w * d * d;
          e_den += w * eps[c] * eps[c];
        }
      }
      store6(P.eps, ns, i, eps);
    } else {
      load6(P.eps, ns, i, eps);
    }
    if (flags & SIC_POST_STRESS) {
      // sig = CT:(eps - eps_rhs)   (compute_stress, MomentumEquation.py:844-866)
      double er[6], d[6], CT[36];
      load6(P.eps_rhs, ns, i, er);
#pragma unroll
      for (int c = 0; c < 6; ++c) d[c] = eps[c] - er[c];
#pragma unroll
      for (int j = 0; j < 36; ++j) CT[j] = P.CT[SIC_CT_INDEX(j, i)];
      ddot66(CT, d, sig);
      store6(P.sig, ns, i, sig);
    } else {
      load6(P.sig, ns, i, sig);
    }
    if (flags & (SIC_POST_INCREMENT | SIC_POST_RATES)) {
      Row row{P.mat_table + (size_t)P.mat_id[i] * P.row_len};
      const double T = P.T[i];
      const double phi1 = dt * theta;
      for (int e = 0; e < P.n_elems; ++e) {
        const sic_elem_t& el = P.elems[e];
        const int off = el.param_off;
        double rate[6];
        if (DESAI && el.kind == SIC_ELEM_DESAI) {
          DesaiP dp = load_desai(row, off);
          double* ds = el.desai;
          double alpha = ds[(size_t)SIC_DS_ALPHA * ns + i];
          const double alpha_0 = ds[(size_t)SIC_DS_ALPHA0 * ns + i];
          if (flags & SIC_POST_INCREMENT) {
            // increment_internal_variables (MaterialProps.py:1129-1158)
            double sk[6], Pv[6], dsig[6];
            load6(P.sig_k, ns, i, sk);
#pragma unroll
            for (int c = 0; c < 6; ++c) {
              Pv[c] = ds[(size_t)(SIC_DS_P + c) * ns + i];
              dsig[c] = sig[c] - sk[c];
            }
            const double r = ds[(size_t)SIC_DS_R * ns + i];
            const double h = ds[(size_t)SIC_DS_H * ns + i];
            const bool h_small = ds[(size_t)SIC_DS_HSMALL * ns + i] != 0.0;
            double d_alpha = (-(r + ddot_sym(Pv, dsig))) / h;
            if (h_small) d_alpha = 0.0;
            alpha = alpha + d_alpha;
            alpha = (alpha < 1.0e-10) ? 1.0e-10 : alpha;
            ds[(size_t)SIC_DS_ALPHA * ns + i] = alpha;
          }
          if (!(flags & SIC_POST_RATES)) continue;
          double fv;
          rate_desai(sig, alpha, alpha_0, dp, rate, fv);
          ds[(size_t)SIC_DS_FVP * ns + i] = fv;
        } else if (EXT && el.kind == SIC_ELEM_MUNSON_DAWSON) {
          MunsonDawsonP mp = load_md(row, off);
          double* ms = el.desai;
          double zeta = ms[(size_t)SIC_MD_ZETA * ns + i];
          if (flags & SIC_POST_INCREMENT) {
            // increment_internal_variables (MaterialProps.py:2081-2102)
            double sk[6], Pv[6], dsig[6];
            load6(P.sig_k, ns, i, sk);
#pragma unroll
            for (int c = 0; c < 6; ++c) {
              Pv[c] = ms[(size_t)(SIC_MD_P + c) * ns + i];
              dsig[c] = sig[c] - sk[c];
            }
            const double r = ms[(size_t)SIC_MD_R * ns + i];
            const double h = ms[(size_t)SIC_MD_H * ns + i];
            const bool h_small = ms[(size_t)SIC_MD_HSMALL * ns + i] != 0.0;
            double d_zeta = (-(r + ddot_sym(Pv, dsig))) / h;
            if (h_small) d_zeta = 0.0;
            zeta = clamp_min(zeta + d_zeta, 0.0);
            ms[(size_t)SIC_MD_ZETA * ns + i] = zeta;
          }
          if (!(flags & SIC_POST_RATES)) continue;
          double Fd, ets;
          rate_munson_dawson(sig, T, zeta, mp, rate, Fd, ets);
          ms[(size_t)SIC_MD_F * ns + i] = Fd;
          ms[(size_t)SIC_MD_ETS * ns + i] = ets;
        } else if (EXT && (el.kind == SIC_ELEM_MOHR_COULOMB || el.kind == SIC_ELEM_MATSUOKA_NAKAI)) {
          if (!(flags & SIC_POST_RATES)) continue;
          double fv;
          if (el.kind == SIC_ELEM_MOHR_COULOMB) rate_mohr_coulomb(sig, load_mc(row, off), rate, fv);
          else rate_matsuoka_nakai(sig, load_mn(row, off), rate, fv);
          el.desai[(size_t)SIC_VP_FVP * ns + i] = fv;
        } else {
          if (!(flags & SIC_POST_RATES)) continue;
          if (el.kind == SIC_ELEM_KELVIN) {
            KelvinP kp{row[off], row[off + 1], row[off + 2], row[off + 3]};
            double g11 = 0.0, g12 = 0.0, g44 = 0.0;
            if (kelvin_phi2 >= 0.0) kelvin_G(kp, kelvin_phi2, g11, g12, g44);
            double eo[6], ro[6];
            load6(el.eps_old, ns, i, eo);
            load6(el.rate_old, ns, i, ro);
            rate_kelvin(sig, kp, g11, g12, g44, eo, ro, phi1, rate);
          } else if (el.kind == SIC_ELEM_DISLOCATION) {
            DislocationP dp{row[off], row[off + 1], row[off + 2]};
            rate_dislocation(sig, T, dp, rate);
          } else {
            PressureSolP pp{row[off], row[off + 1], row[off + 2]};
            rate_pressure_solution(sig, T, pp, rate);
          }
        }
        store6(el.rate, ns, i, rate);
      }
    }
  }
  if (flags & SIC_POST_ERROR) {
    // deterministic block reduction; partials summed by k_post_err_final
    __shared__ double sh[2][SIC_CELL_THREADS / 32];
    e_num = warp_sum(e_num);
    e_den = warp_sum(e_den);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = e_num; sh[1][w] = e_den; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int k = 0; k < SIC_CELL_THREADS / 32; ++k) { a += sh[0][k]; b += sh[1][k]; }
      err_scratch[2 * blockIdx.x] = a;
      err_scratch[2 * blockIdx.x + 1] = b;
    }
  }
}

__global__ void k_post_err_final(const double* __restrict__ scratch, int n_blocks, double* __restrict__ out) {
  __shared__ double sh[2][32];
  double a = 0.0, b = 0.0;
  for (int k = threadIdx.x; k < n_blocks; k += blockDim.x) { a += scratch[2 * k]; b += scratch[2 * k + 1]; }
  a = warp_sum(a); b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = a; sh[1][w] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double x = 0.0, y = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { x += sh[0][k]; y += sh[1][k]; }
    out[0] = x; out[1] = y;
  }
}

// =============================================================================================
// commit of a converged step (Simulators.py:509-517)
// =============================================================================================
template <int SET>
__global__ void __launch_bounds__(SIC_CELL_THREADS) k_commit(sic_problem_t P, double dt, double theta) {
  constexpr bool DESAI = SET >= 1, EXT = SET >= 2;   // element set of the instantiation (see sic_element_set)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_cells) return;
  const int ns = P.cell_stride;
  const double phi2 = dt * (1.0 - theta);
  Row row{P.mat_table + (size_t)P.mat_id[i] * P.row_len};
  const double T = P.T[i];
  double sig[6], sk[6], dsig[6];
  load6(P.sig, ns, i, sig);
  load6(P.sig_k, ns, i, sk);
#pragma unroll
  for (int c = 0; c < 6; ++c) dsig[c] = sig[c] - sk[c];

  for (int e = 0; e < P.n_elems; ++e) {
    const sic_elem_t& el = P.elems[e];
    const int off = el.param_off;
    double Gd[6] = {0, 0, 0, 0, 0, 0};  // G_i : (sig - sig_k), columns accumulated in order
    double Bv[6] = {0, 0, 0, 0, 0, 0};
    auto acc = [&](int k, const double* col) {
      double dk = 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) dk = (c == k) ? dsig[c] : dk;
#pragma unroll
      for (int r = 0; r < 6; ++r) Gd[r] = (k == 0) ? col[r] * dk : Gd[r] + col[r] * dk;
    };
    if (el.kind == SIC_ELEM_KELVIN) {
      KelvinP kp{row[off], row[off + 1], row[off + 2], row[off + 3]};
      double g11, g12, g44;
      kelvin_G(kp, phi2, g11, g12, g44);
      ddot_iso(g11, g12, g44, dsig, Gd);
    } else if (el.kind == SIC_ELEM_DISLOCATION) {
      DislocationP dp{row[off], row[off + 1], row[off + 2]};
      fd_columns<SET == 0 && SIC_TAN_UNROLL>([&](const double* s, double* r) { rate_dislocation(s, T, dp, r); }, sk, acc);
    } else if (el.kind == SIC_ELEM_PRESSURE_SOL) {
      PressureSolP pp{row[off], row[off + 1], row[off + 2]};
      fd_columns<SET == 0 && SIC_TAN_UNROLL>([&](const double* s, double* r) { rate_pressure_solution(s, T, pp, r); }, sk, acc);
    } else if (DESAI && el.kind == SIC_ELEM_DESAI) {
      DesaiP dp = load_desai(row, off);
      double* ds = el.desai;
      DesaiLin L;
      L.r = ds[(size_t)SIC_DS_R * ns + i];
      L.h = ds[(size_t)SIC_DS_H * ns + i];
      L.h_small = ds[(size_t)SIC_DS_HSMALL * ns + i] != 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        L.P[c] = ds[(size_t)(SIC_DS_P + c) * ns + i];
        L.Q[c] = ds[(size_t)(SIC_DS_Q + c) * ns + i];
      }
      const double alpha_k = ds[(size_t)SIC_DS_ALPHA_K * ns + i];
      const double alpha_0 = ds[(size_t)SIC_DS_ALPHA0 * ns + i];
      if (!L.h_small) {
        double rh = L.r / L.h;
#pragma unroll
        for (int c = 0; c < 6; ++c) Bv[c] = rh * L.Q[c];
      }
      fd_columns([&](const double* s, double* r) { double fv; rate_desai(s, alpha_k, alpha_0, dp, r, fv); }, sk,
                 [&](int k, const double* col) {
                   double gc[6];
#pragma unroll
                   for (int r = 0; r < 6; ++r) gc[r] = col[r] - desai_Hh(L, r, k);
                   acc(k, gc);
                 });
      // update_internal_variables (MaterialProps.py:1119-1127)
      ds[(size_t)SIC_DS_QSI_OLD * ns + i] = ds[(size_t)SIC_DS_QSI * ns + i];
    }
    if constexpr (EXT) {
      if (el.kind == SIC_ELEM_MUNSON_DAWSON) {
        MunsonDawsonP mp = load_md(row, off);
        double* ms = el.desai;
        DesaiLin L;
        L.r = ms[(size_t)SIC_MD_R * ns + i];
        L.h = ms[(size_t)SIC_MD_H * ns + i];
        L.h_small = ms[(size_t)SIC_MD_HSMALL * ns + i] != 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          L.P[c] = ms[(size_t)(SIC_MD_P + c) * ns + i];
          L.Q[c] = ms[(size_t)(SIC_MD_Q + c) * ns + i];
        }
        const double zeta_k = ms[(size_t)SIC_MD_ZETA_K * ns + i];
        if (!L.h_small) {
          double rh = L.r / L.h;
#pragma unroll
          for (int c = 0; c < 6; ++c) Bv[c] = rh * L.Q[c];
        }
        fd_columns([&](const double* s, double* r) { double f, e; rate_munson_dawson(s, T, zeta_k, mp, r, f, e); }, sk,
                   [&](int k, const double* col) {
                     double gc[6];
#pragma unroll
                     for (int r = 0; r < 6; ++r) gc[r] = col[r] - desai_Hh(L, r, k);
                     acc(k, gc);
                   });
        // update_internal_variables (MaterialProps.py:2077-2079)
        ms[(size_t)SIC_MD_ZETA_OLD * ns + i] = ms[(size_t)SIC_MD_ZETA * ns + i];
      } else if (el.kind == SIC_ELEM_MOHR_COULOMB) {
        MohrCoulombP cp = load_mc(row, off);
        fd_columns([&](const double* s, double* r) { double f; rate_mohr_coulomb(s, cp, r, f); }, sk, acc);
      } else if (el.kind == SIC_ELEM_MATSUOKA_NAKAI) {
        MatsuokaNakaiP np_ = load_mn(row, off);
        fd_columns([&](const double* s, double* r) { double f; rate_matsuoka_nakai(s, np_, r, f); }, sk, acc);
      }
    }
    // update_eps_ne_rate_old (:630-638)
    double rate[6];
    load6(el.rate, ns, i, rate);
    store6(el.rate_old, ns, i, rate);
    // update_eps_ne_old (:607-628): eps_old = eps_k + phi2 G_i:(sig - sig_k) - phi2 B_i
    double ek[6], eo[6];
    load6(el.eps_k, ns, i, ek);
#pragma unroll
    for (int c = 0; c < 6; ++c) eo[c] = (ek[c] + phi2 * Gd[c]) - phi2 * Bv[c];
    store6(el.eps_old, ns, i, eo);
  }
}

__global__ void __launch_bounds__(SIC_CELL_THREADS) k_commit_rates(sic_problem_t P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_cells) return;
  const int ns = P.cell_stride;
  for (int e = 0; e < P.n_elems; ++e) {
    double rate[6];
    load6(P.elems[e].rate, ns, i, rate);
    store6(P.elems[e].rate_old, ns, i, rate);
  }
}

// ViscoplasticDesai.compute_initial_hardening (MaterialProps.py:1248-1288)
__global__ void __launch_bounds__(SIC_CELL_THREADS) k_desai_init(sic_problem_t P, int e, double Fvp_0,
                                                                int32_t* n_clamped) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_cells) return;
  const int ns = P.cell_stride;
  Row row{P.mat_table + (size_t)P.mat_id[i] * P.row_len};
  const sic_elem_t& el = P.elems[e];
  DesaiP p = load_desai(row, el.param_off);
  double sig[6];
  load6(P.sig, ns, i, sig);
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
  J2 = (J2 < 1.0e-6) ? 1.0e-6 : J2;
  double Sr = low_J2 ? 0.0 : (-(J3 * SIC_SQRT27)) / (2.0 * sic_pow(J2, 1.5));
  double I1s = I1 + p.sigma_t;
  double F2 = sic_exp(p.beta_1 * I1s) - p.beta * Sr;
  F2 = (F2 < 1.0e-6) ? 1.0e-6 : F2;
  double a0 = p.gamma * sic_pow(I1s, 2.0 - p.n) + ((Fvp_0 - J2) * sic_pow(I1s, -p.n)) * sic_pow(F2, -p.m);
  if (a0 <= 1.0e-6 && n_clamped) atomicAdd(n_clamped, 1);
  a0 = (a0 < 1.0e-6) ? 1.0e-6 : a0;
  double* ds = el.desai;
  ds[(size_t)SIC_DS_ALPHA0 * ns + i] = a0;
  ds[(size_t)SIC_DS_ALPHA * ns + i] = a0;
  double F1 = a0 * sic_pow(I1s, p.n) - p.gamma * (I1s * I1s);
  ds[(size_t)SIC_DS_FVP * ns + i] = J2 + F1 * sic_pow(F2, p.m);
}

}  // namespace sic

// =============================================================================================
// C ABI
// =============================================================================================
using namespace sic;

static int check_problem(const sic_problem_t* p) {
  if (!p) return sic_fail("null problem");
  if (p->abi_version != SIC_ABI_VERSION) return sic_fail("sic_problem_t.abi_version mismatch");
  if (p->n_cells < 0 || p->cell_stride < p->n_cells) return sic_fail("bad n_cells / cell_stride");
  if (p->n_elems < 0 || p->n_elems > SIC_MAX_ELEMS) return sic_fail("too many non-elastic elements");
  if (p->n_thermo < 0 || p->n_thermo > SIC_MAX_THERMO) return sic_fail("too many thermoelastic elements");
  for (int e = 0; e < p->n_elems; ++e) {
    int k = p->elems[e].kind;
    if (k < SIC_ELEM_KELVIN || k > SIC_ELEM_MATSUOKA_NAKAI) return sic_fail("unknown element kind");
    if (k >= SIC_ELEM_DESAI && !p->elems[e].desai) return sic_fail("element with internal state but without its state block");
  }
  return 0;
}

static inline int cell_blocks(int n) { return (n + SIC_CELL_THREADS - 1) / SIC_CELL_THREADS; }

// Which instantiation of k_tangent / k_post / k_commit a material runs: 0 = Kelvin / DislocationCreep /
// PressureSolutionCreep only (BASELINE configs 1-2 and the bench), 1 = + ViscoplasticDesai (configs 3-4), 2 = + the
// SURVEY 8f elements.  The smaller sets are the same code with the unused branches compiled out: fewer registers
// (the Desai linearisation is what pushes the full kernels to 255 registers with spills), hence more resident warps.
static inline int sic_element_set(const sic_problem_t* p) {
  int set = 0;
  for (int e = 0; e < p->n_elems; ++e) {
    if (p->elems[e].kind > SIC_ELEM_DESAI) return 2;
    if (p->elems[e].kind == SIC_ELEM_DESAI) set = 1;
  }
  return set;
}
#define SIC_LAUNCH_SET(kernel, p, grid, stream, ...)                                                     \
  do {                                                                                                   \
    switch (sic_element_set(p)) {                                                                        \
      case 0: kernel<0><<<grid, SIC_CELL_THREADS, 0, (cudaStream_t)stream>>>(__VA_ARGS__); break;        \
      case 1: kernel<1><<<grid, SIC_CELL_THREADS, 0, (cudaStream_t)stream>>>(__VA_ARGS__); break;        \
      default: kernel<2><<<grid, SIC_CELL_THREADS, 0, (cudaStream_t)stream>>>(__VA_ARGS__); break;       \
    }                                                                                                    \
  } while (0)

extern "C" int sic_tangent(const sic_problem_t* p, double dt, double theta, void* stream) {
  if (int rc = check_problem(p)) return rc;
  if (p->n_cells == 0) return 0;
  SIC_LAUNCH_SET(k_tangent, p, cell_blocks(p->n_cells), stream, *p, dt, theta);
  return sic_check_launch("k_tangent");
}

extern "C" int sic_elastic_tangent(const sic_problem_t* p, void* stream) {
  if (int rc = check_problem(p)) return rc;
  if (p->n_cells == 0) return 0;
  k_elastic_tangent<<<cell_blocks(p->n_cells), SIC_CELL_THREADS, 0, (cudaStream_t)stream>>>(*p);
  return sic_check_launch("k_elastic_tangent");
}

extern "C" int sic_post_blocks(int n_cells) { return cell_blocks(n_cells); }

extern "C" int sic_post(const sic_problem_t* p, const double* u, double dt, double theta, double kelvin_phi2,
                        int flags, double* err_out, double* err_scratch, void* stream) {
  if (int rc = check_problem(p)) return rc;
  if ((flags & SIC_POST_STRAIN) && !u) return sic_fail("sic_post: SIC_POST_STRAIN needs u");
  if ((flags & SIC_POST_ERROR) && (!(flags & SIC_POST_STRAIN) || !err_out || !err_scratch))
    return sic_fail("sic_post: SIC_POST_ERROR needs SIC_POST_STRAIN, err_out and err_scratch");
  if (p->n_cells == 0) return 0;
  const int nb = cell_blocks(p->n_cells);
  SIC_LAUNCH_SET(k_post, p, nb, stream, *p, u, dt, theta, kelvin_phi2, flags, err_scratch);
  if (int rc = sic_check_launch("k_post")) return rc;
  if (flags & SIC_POST_ERROR) {
    k_post_err_final<<<1, 1024, 0, (cudaStream_t)stream>>>(err_scratch, nb, err_out);
    return sic_check_launch("k_post_err_final");
  }
  return 0;
}

extern "C" int sic_commit(const sic_problem_t* p, double dt, double theta, void* stream) {
  if (int rc = check_problem(p)) return rc;
  if (p->n_cells == 0 || p->n_elems == 0) return 0;
  SIC_LAUNCH_SET(k_commit, p, cell_blocks(p->n_cells), stream, *p, dt, theta);
  return sic_check_launch("k_commit");
}

extern "C" int sic_commit_rates(const sic_problem_t* p, void* stream) {
  if (int rc = check_problem(p)) return rc;
  if (p->n_cells == 0 || p->n_elems == 0) return 0;
  k_commit_rates<<<cell_blocks(p->n_cells), SIC_CELL_THREADS, 0, (cudaStream_t)stream>>>(*p);
  return sic_check_launch("k_commit_rates");
}

extern "C" int sic_desai_initial_hardening(const sic_problem_t* p, int elem, double Fvp_0, int32_t* n_clamped,
                                           void* stream) {
  if (int rc = check_problem(p)) return rc;
  if (elem < 0 || elem >= p->n_elems || p->elems[elem].kind != SIC_ELEM_DESAI)
    return sic_fail("sic_desai_initial_hardening: element is not a Desai element");
  if (p->n_cells == 0) return 0;
  k_desai_init<<<cell_blocks(p->n_cells), SIC_CELL_THREADS, 0, (cudaStream_t)stream>>>(*p, elem, Fvp_0, n_clamped);
  return sic_check_launch("k_desai_init");
}
