// sic_math.h — bit-reproducible double-precision exp / log / pow for host and device.
//
// Why this exists: the reference evaluates its creep tangents by finite differences with
// ABSOLUTE steps of 1e-2 Pa on stresses of ~1e7 Pa (safeincave/MaterialProps.py:640-675,
// 1459-1488).  A one-ulp difference between two libm implementations of pow()/exp() is
// amplified by sigma/(2*eps_FD) ~ 5e8 in G and C_T.  CUDA's pow/exp (<= 2 ulp) and glibc's
// (< 1 ulp) do not agree bit for bit, so a kernel using the built-ins can only match a CPU
// oracle to ~1e-7 in C_T.  Every routine below is built only from IEEE-754 correctly rounded
// primitives (+ - * / fma rint) and integer bit operations, with the SAME operation order on
// both sides, so the CUDA kernels (nvcc, sm_100a) and the CPU oracle shim (gcc) produce
// IDENTICAL bits.  Accuracy: < 1 ulp (checked against mpmath in tests/test_sic_math.py).
//
// All fused multiply-adds are explicit fma() calls; translation units that include this
// header for parity-critical work are compiled with contraction OFF
// (nvcc -fmad=false, gcc -ffp-contract=off).
#ifndef SIC_MATH_H_
#define SIC_MATH_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SIC_HD __host__ __device__ __forceinline__
#else
#define SIC_HD static inline
#endif

SIC_HD uint64_t sic_d2u(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
SIC_HD double sic_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x; memcpy(&x, &u, 8); return x;
#endif
}

#define SIC_LN2_HI 6.93147180369123816490e-01  /* 0x3FE62E42FEE00000: 21 trailing zero bits */
#define SIC_LN2_LO 1.90821492927058770002e-10  /* ln2 - SIC_LN2_HI */
#define SIC_INV_LN2 1.44269504088896338700e+00

// error-free transformations
SIC_HD void sic_two_sum(double a, double b, double* s, double* e) {
  double t = a + b;
  double bb = t - a;
  *e = (a - (t - bb)) + (b - bb);
  *s = t;
}
SIC_HD void sic_fast_two_sum(double a, double b, double* s, double* e) {  // |a| >= |b|
  double t = a + b;
  *e = b - (t - a);
  *s = t;
}

// exp(h + l) for a double-double argument (|l| << |h|), result rounded to double.
SIC_HD double sic_exp_dd(double h, double l) {
  if (h != h) return h;
  if (h > 709.782712893384) return (double)INFINITY;
  if (h < -745.2) return 0.0;
  double n = rint(h * SIC_INV_LN2);
  double r = fma(-n, SIC_LN2_HI, h);    // exact: n*LN2_HI is exact, difference representable
  double rl = fma(-n, SIC_LN2_LO, l);
  double rh, re;
  sic_two_sum(r, rl, &rh, &re);
  // H(r) = sum_{k=2}^{14} r^(k-2)/k!
  double p = 1.1470745597729725e-11;            // 1/14!
  p = fma(p, rh, 1.6059043836821613e-10);       // 1/13!
  p = fma(p, rh, 2.08767569878681e-09);         // 1/12!
  p = fma(p, rh, 2.505210838544172e-08);        // 1/11!
  p = fma(p, rh, 2.755731922398589e-07);        // 1/10!
  p = fma(p, rh, 2.7557319223985893e-06);       // 1/9!
  p = fma(p, rh, 2.48015873015873e-05);         // 1/8!
  p = fma(p, rh, 1.984126984126984e-04);        // 1/7!
  p = fma(p, rh, 1.388888888888889e-03);        // 1/6!
  p = fma(p, rh, 8.333333333333333e-03);        // 1/5!
  p = fma(p, rh, 4.1666666666666664e-02);       // 1/4!
  p = fma(p, rh, 1.6666666666666666e-01);       // 1/3!
  p = fma(p, rh, 0.5);                          // 1/2!
  double q = (rh * rh) * p;
  double lowpart = q + fma(re, rh, re);         // + re*(1+rh)
  double y = 1.0 + (rh + lowpart);
  // scale by 2^n in two steps (n in [-1075, 1024])
  int ni = (int)n;
  int n1 = ni / 2;
  int n2 = ni - n1;
  double s1 = sic_u2d((uint64_t)(1023 + n1) << 52);
  double s2 = sic_u2d((uint64_t)(1023 + n2) << 52);
  return (y * s1) * s2;
}

SIC_HD double sic_exp(double x) { return sic_exp_dd(x, 0.0); }

// log(x) as a double-double (hi, lo) for finite x > 0.
SIC_HD void sic_log_dd(double x, double* hi, double* lo) {
  int k = 0;
  uint64_t u = sic_d2u(x);
  if ((u >> 52) == 0) {  // subnormal: scale by 2^54
    x = x * 18014398509481984.0;
    u = sic_d2u(x);
    k = -54;
  }
  k += (int)(u >> 52) - 1023;
  u = (u & 0x000FFFFFFFFFFFFFULL) | 0x3FF0000000000000ULL;  // m in [1,2)
  double m = sic_u2d(u);
  if (m > 1.4142135623730951) { m = m * 0.5; k += 1; }       // m in (sqrt(1/2), sqrt(2)]
  double f = m - 1.0;  // exact
  double ah, al;
  sic_two_sum(m, 1.0, &ah, &al);
  double sh = f / ah;
  double rem = fma(-sh, ah, f);
  rem = fma(-sh, al, rem);
  double sl = rem / ah;
  double s2 = sh * sh;
  // atanh(s) - s = s^3 * (1/3 + s^2/5 + ... + s^24/27)
  double p = 3.7037037037037035e-02;          // 1/27
  p = fma(p, s2, 4.0e-02);                    // 1/25
  p = fma(p, s2, 4.3478260869565216e-02);     // 1/23
  p = fma(p, s2, 4.7619047619047616e-02);     // 1/21
  p = fma(p, s2, 5.2631578947368418e-02);     // 1/19
  p = fma(p, s2, 5.8823529411764705e-02);     // 1/17
  p = fma(p, s2, 6.6666666666666666e-02);     // 1/15
  p = fma(p, s2, 7.6923076923076927e-02);     // 1/13
  p = fma(p, s2, 9.0909090909090912e-02);     // 1/11
  p = fma(p, s2, 1.1111111111111111e-01);     // 1/9
  p = fma(p, s2, 1.4285714285714285e-01);     // 1/7
  p = fma(p, s2, 2.0e-01);                    // 1/5
  p = fma(p, s2, 3.3333333333333331e-01);     // 1/3
  double t = (sh * s2) * p;
  double small = 2.0 * (sl + t);
  double kd = (double)k;
  double H, L;
  sic_two_sum(kd * SIC_LN2_HI, 2.0 * sh, &H, &L);
  L = L + fma(kd, SIC_LN2_LO, small);
  sic_fast_two_sum(H, L, hi, lo);
}

SIC_HD double sic_log(double x) {
  if (x != x) return x;
  if (x < 0.0) return (double)NAN;
  if (x == 0.0) return -(double)INFINITY;
  if (x == (double)INFINITY) return x;
  double h, l;
  sic_log_dd(x, &h, &l);
  return h;
}

// pow(x, y) with C99 semantics for the cases the constitutive laws can reach
// (negative base with integral exponent keeps its sign; negative base with a
// fractional exponent is NaN, as torch/numpy give).
// log10(x) = log(x) * (1/ln 10): one extra rounding on top of sic_log (MunsonDawsonCreep's Delta, MaterialProps.py:2153)
SIC_HD double sic_log10(double x) { return sic_log(x) * 0.43429448190325182765; }

SIC_HD double sic_pow(double x, double y) {
  if (y == 0.0) return 1.0;
  // Exact small-integer exponents: one correctly rounded product IS the correctly rounded power, for every x
  // (incl. negative, infinite and NaN arguments), and it is what the creep laws of every BASELINE configuration ask
  // for 12 times per cell and tangent: DislocationCreep's q^(n-1) with n = 3.
  if (y == 2.0) return x * x;
  if (y == 1.0) return x;
  if (x == 1.0) return 1.0;
  if (x != x || y != y) return x + y;
  double sign = 1.0;
  if (x < 0.0) {
    double yi = rint(y);
    if (yi != y || fabs(y) > 9007199254740992.0) {
      if (fabs(y) == (double)INFINITY) {
        x = -x;
      } else if (yi != y) {
        return (double)NAN;
      } else {
        x = -x;  // huge even integer
      }
    } else {
      double half = yi * 0.5;
      if (rint(half) != half) sign = -1.0;
      x = -x;
    }
  }
  if (x == 0.0) return (y > 0.0) ? 0.0 * sign : (double)INFINITY;
  if (x == (double)INFINITY) return (y > 0.0) ? sign * (double)INFINITY : 0.0;
  if (fabs(y) == (double)INFINITY) {
    if (x == 1.0) return 1.0;
    return ((x > 1.0) == (y > 0.0)) ? (double)INFINITY : 0.0;
  }
  double lh, ll;
  sic_log_dd(x, &lh, &ll);
  double ph = y * lh;
  double pl = fma(y, lh, -ph);
  pl = fma(y, ll, pl);
  if (ph > 1.0e4) return sign * (double)INFINITY;
  if (ph < -1.0e4) return sign * 0.0;
  return sign * sic_exp_dd(ph, pl);
}

#endif  // SIC_MATH_H_
