/* CPU restatement of exp_fast / exp_clamped (csrc/gram.cuh) with the same operation order and fma contraction points,
 * so that its accuracy can be pinned without a GPU (tests/test_fast_exp.py compiles this with gcc -mfma). */
#include <math.h>
#include <stdint.h>
#include <string.h>

static const double EXP_C[18] = {
    6755399441055744.0, 1.4426950408889634, -6.93147180369123816490e-01, -1.90821492927058770002e-10,
    1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0,
    1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0};

static double exp_fast(double x) {
  const double t = fma(x, EXP_C[1], EXP_C[0]);
  int64_t bits;
  memcpy(&bits, &t, 8);
  const int n = (int)(bits & 0xffffffff);
  const double nd = t - EXP_C[0];
  double r = fma(nd, EXP_C[2], x);
  r = fma(nd, EXP_C[3], r);
  double p = EXP_C[4];
  for (int k = 5; k < 18; ++k) p = fma(p, r, EXP_C[k]);
  memcpy(&bits, &p, 8);
  bits += ((int64_t)n) << 52;
  memcpy(&p, &bits, 8);
  return p;
}

double exp_clamped_model(double x) {
  const double e = exp_fast(x);
  uint64_t bits;
  memcpy(&bits, &x, 8);
  const int tiny = (uint32_t)(bits >> 32) >= 0xC0862000u; /* x <= -708.0 (or NaN with the sign bit): integer compare as in the kernel */
  return tiny ? 0.0 : e;
}

/* max relative error against long double expl over n points uniformly spread on [lo, hi] */
double exp_model_max_rel_err(double lo, double hi, int64_t n) {
  double worst = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const double x = lo + (hi - lo) * ((double)i + 0.37) / (double)n;
    const long double ref = expl((long double)x);
    const double rel = fabs((double)(((long double)exp_clamped_model(x) - ref) / ref));
    if (rel > worst) worst = rel;
  }
  return worst;
}
