/* Array front-ends for safeincave_b200/csrc/sic_math.h, for the numpy oracle.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md).  The oracle restates the reference's
 * constitutive algorithm in numpy; the only thing it shares with the product is this
 * elementary-function header, so that finite-difference tangents (amplification ~5e8,
 * safeincave/MaterialProps.py:640-675) can be compared bit for bit.
 * Build: gcc -O2 -ffp-contract=off -mfma -shared -fPIC (see oracle/Makefile).
 */
#include "../safeincave_b200/csrc/sic_math.h"

void sic_pow_array(const double* x, const double* y, double* out, long n) {
  for (long i = 0; i < n; ++i) out[i] = sic_pow(x[i], y[i]);
}
void sic_exp_array(const double* x, double* out, long n) {
  for (long i = 0; i < n; ++i) out[i] = sic_exp(x[i]);
}
void sic_log_array(const double* x, double* out, long n) {
  for (long i = 0; i < n; ++i) out[i] = sic_log(x[i]);
}
void sic_log10_array(const double* x, double* out, long n) {
  for (long i = 0; i < n; ++i) out[i] = sic_log10(x[i]);
}
