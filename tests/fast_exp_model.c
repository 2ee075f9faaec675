/* CPU restatement of exp2_tab (csrc/gram.cuh), the whole RBF epilogue: 2^(y/256) with a 256-entry table of 2^(j/256), exact
 * reduction by the magic-number add, e^z - 1 = z + z^2 (b0 + b1 z + b2 z^2) on |z| <= ln2/512, exponent-field add and the integer
 * clamp of N — same operation order and fma contraction points, so that its accuracy can be pinned without a GPU
 * (tests/test_fast_exp.py compiles this with gcc -mfma). The table is built as in ctx_create (plb200_api.cu): exp2l rounded to double. */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define TAB_BITS 8
#define TAB_N (1 << TAB_BITS)
static double TAB[TAB_N];
static int tab_ready = 0;
static const double EXP2_C[5] = {6755399441055744.0, 0.0027076061740622863, 3.6655655969101062e-06, 3.308302907918888e-09, 2.239395293483284e-12};

static void tab_init(void) {
  if (tab_ready) return;
  for (int j = 0; j < TAB_N; ++j) TAB[j] = (double)exp2l((long double)j / (long double)TAB_N);
  tab_ready = 1;
}

/* 2^(y/256) */
double exp2_tab_model(double y) {
  tab_init();
  const double t = y + EXP2_C[0];
  int64_t bits;
  memcpy(&bits, &t, 8);
  int N = (int)(bits & 0xffffffff);
  if (N < -1022 * TAB_N) N = -1022 * TAB_N; /* keeps the exponent field in range: results below 2^-1022 come out as <= 2^-1021 */
  const double r = y - (t - EXP2_C[0]);
  const double T = TAB[N & (TAB_N - 1)];
  double q = fma(r, EXP2_C[4], EXP2_C[3]);
  q = fma(q, r, EXP2_C[2]);
  q = fma(q, r, EXP2_C[1]);
  double e = fma(T * r, q, T);
  memcpy(&bits, &e, 8);
  bits += ((int64_t)(N >> TAB_BITS)) << 52;
  memcpy(&e, &bits, 8);
  return e;
}

/* the scale the prelude folds into the operands: y = 256 log2(e) x */
double exp2_scale(void) { return 256.0 * 1.4426950408889634074; }

/* max relative error against long double exp2l over n points y uniformly spread on [lo, hi] (y in units of 1/256 octave) */
double exp2_model_max_rel_err(double lo, double hi, int64_t n) {
  double worst = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const double y = lo + (hi - lo) * ((double)i + 0.37) / (double)n;
    const long double ref = exp2l((long double)y / (long double)TAB_N);
    const double rel = fabs((double)(((long double)exp2_tab_model(y) - ref) / ref));
    if (rel > worst) worst = rel;
  }
  return worst;
}
