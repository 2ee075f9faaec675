/*
 * c_abi_driver.c — exercises libplb200.so exactly the way the Julia shim's `ccall`s do, with no Python in between:
 * plain C, dlopen, the same argument lists (PotentialLearningB200/src/PotentialLearningB200.jl), buffers laid out as Julia lays
 * them out (column-major; Vector{Vector} packed with reduce(hcat, F)), results compared with the CPU oracle (libploracle.so,
 * test infrastructure) in Julia's own indexing K[i + j*n].
 *
 *   gcc -O1 -o c_abi_driver tests/c_abi_driver.c -ldl -lm
 *   ./c_abi_driver <libplb200.so> <libploracle.so> [ndev]      -> prints "OK <checks>" and exits 0, or the first failure and exits 1
 *
 * Run from tests/test_c_abi_driver.py (-m gpu). Tolerance 1e-11 relative (contract 1e-10).
 */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define RTOL 1e-11
#define PL_KERNEL_DOT 1
#define PL_KERNEL_RBF 2
#define PL_FILL_JULIA_U 2

typedef int (*fn_init_all)(int, const int*);
typedef void (*fn_finalize)(void);
typedef const char* (*fn_last_error)(void);
typedef int (*fn_gram_dot)(const double*, int64_t, int64_t, int64_t, int, double*, int64_t, int);
typedef int (*fn_gram_rbf)(const double*, int64_t, int64_t, int64_t, int, const double*, double, double, double*, int64_t, int);
typedef int (*fn_gram_cross)(int, const double*, int64_t, int64_t, const double*, int64_t, int64_t, int64_t, int, int, const double*, double, double,
                             double*, int64_t);
typedef int (*fn_normal_eq)(const double*, int64_t, int64_t, int64_t, const double*, int64_t, double, double, const double*, int, double, double*,
                            double*);
typedef int (*fn_normal_eq_blocks)(const double* const*, const int64_t*, int64_t, int64_t, const double*, int64_t, double, double, const double*, int,
                                   double, int64_t, double*, double*);
typedef int (*fn_cov)(const double*, int64_t, int64_t, int64_t, double*, int, int, double*);
typedef int (*fn_device_count)(void);

typedef int (*orc_sym)(int, const double*, int64_t, int64_t, int64_t, int, int, const double*, double, double, double*, int64_t, int64_t, int64_t);
typedef int (*orc_cross)(int, const double*, int64_t, int64_t, const double*, int64_t, int64_t, int64_t, int, int, const double*, double, double, double*,
                         int64_t);
typedef int (*orc_atqa_t)(const double*, int64_t, int64_t, int64_t, const double*, const double*, double*, double*);
typedef int (*orc_outer_t)(const double*, int64_t, int64_t, int64_t, const double*, const double*, const double*, double*, double*);

static fn_last_error pl_last_error;
static int checks = 0;

/* xorshift64* + Box-Muller: deterministic inputs without any library */
static uint64_t rs = 0x9E3779B97F4A7C15ull;
static double urand(void) {
  rs ^= rs >> 12;
  rs ^= rs << 25;
  rs ^= rs >> 27;
  return (double)((rs * 0x2545F4914F6CDD1Dull) >> 11) / 9007199254740992.0;
}
static double nrand(void) {
  double u = urand(), v = urand();
  if (u < 1e-300) u = 1e-300;
  return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v);
}

static void* sym(void* h, const char* name) {
  void* p = dlsym(h, name);
  if (!p) {
    fprintf(stderr, "missing symbol %s\n", name);
    exit(1);
  }
  return p;
}
#define CALL(expr)                                                              \
  do {                                                                          \
    int rc_ = (expr);                                                           \
    if (rc_ != 0) {                                                             \
      fprintf(stderr, "%s failed (code %d): %s\n", #expr, rc_, pl_last_error()); \
      exit(1);                                                                  \
    }                                                                           \
  } while (0)

static void expect_close(const char* what, double got, double ref, long i, long j) {
  const double err = fabs(got - ref) / (fabs(ref) > 0 ? fabs(ref) : 1.0);
  if (!(err < RTOL)) {
    fprintf(stderr, "FAILED %s at (%ld, %ld): got %.17g, reference %.17g (rel %.3g)\n", what, i, j, got, ref, err);
    exit(1);
  }
}
static void expect_norm_close(const char* what, const double* got, const double* ref, long count) {
  double scale = 0, worst = 0;
  for (long e = 0; e < count; ++e)
    if (fabs(ref[e]) > scale) scale = fabs(ref[e]);
  for (long e = 0; e < count; ++e) {
    double err = fabs(got[e] - ref[e]) / scale;
    if (fabs(ref[e]) > 1e-3 * scale) {
      const double r = fabs(got[e] - ref[e]) / fabs(ref[e]);
      if (r > err) err = r;
    }
    if (err > worst) worst = err;
  }
  if (!(worst < RTOL)) {
    fprintf(stderr, "FAILED %s: worst error %.3g\n", what, worst);
    exit(1);
  }
  ++checks;
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s libplb200.so libploracle.so [ndev]\n", argv[0]);
    return 2;
  }
  void* hp = dlopen(argv[1], RTLD_NOW);
  void* ho = dlopen(argv[2], RTLD_NOW);
  if (!hp || !ho) {
    fprintf(stderr, "dlopen failed: %s\n", dlerror());
    return 1;
  }
  const int ndev = argc > 3 ? atoi(argv[3]) : 1;
  fn_init_all pl_init_all = (fn_init_all)sym(hp, "pl_init_all");
  fn_finalize pl_finalize = (fn_finalize)sym(hp, "pl_finalize");
  pl_last_error = (fn_last_error)sym(hp, "pl_last_error");
  fn_gram_dot pl_gram_dot = (fn_gram_dot)sym(hp, "pl_gram_dot");
  fn_gram_rbf pl_gram_rbf = (fn_gram_rbf)sym(hp, "pl_gram_rbf");
  fn_gram_cross pl_gram_cross = (fn_gram_cross)sym(hp, "pl_gram_cross");
  fn_normal_eq pl_normal_eq = (fn_normal_eq)sym(hp, "pl_normal_eq");
  fn_normal_eq_blocks pl_normal_eq_blocks = (fn_normal_eq_blocks)sym(hp, "pl_normal_eq_blocks");
  fn_cov pl_cov = (fn_cov)sym(hp, "pl_cov");
  fn_device_count pl_device_count = (fn_device_count)sym(hp, "pl_device_count");
  orc_sym orc_kernel_matrix_sym = (orc_sym)sym(ho, "orc_kernel_matrix_sym");
  orc_cross orc_kernel_matrix_cross = (orc_cross)sym(ho, "orc_kernel_matrix_cross");
  orc_atqa_t orc_atqa = (orc_atqa_t)sym(ho, "orc_atqa");
  orc_outer_t orc_outer_sum = (orc_outer_t)sym(ho, "orc_outer_sum");

  CALL(pl_init_all(ndev, NULL)); /* __init__() of the shim */
  if (pl_device_count() != (ndev > 0 ? ndev : pl_device_count())) return 1;

  /* ---- KernelMatrix(F, k): F::Vector{Vector{Float64}} -> X = reduce(hcat, F) (d x n column-major), K = zeros(n, n) ---- */
  {
    const int64_t n = 333, d = 26; /* DPP-ACE-Si descriptor length */
    double* X = (double*)malloc(sizeof(double) * n * d);
    double* mu = (double*)malloc(sizeof(double) * d);
    for (int64_t k = 0; k < d; ++k) mu[k] = 2.0 * nrand();
    for (int64_t i = 0; i < n; ++i)
      for (int64_t k = 0; k < d; ++k) X[i * d + k] = mu[k] + 0.3 * nrand(); /* column i of the Julia matrix */
    double* K = (double*)calloc(n * n, sizeof(double));
    double* Kref = (double*)calloc(n * n, sizeof(double));
    /* ccall((:pl_gram_dot, lib), Cint, (Ptr{Cdouble}, Int64, Int64, Int64, Cint, Ptr{Cdouble}, Int64, Cint), X, n, d, d, Cint(k.α), K, n, PL_FILL_JULIA_U) */
    CALL(pl_gram_dot(X, n, d, d, 2, K, n, PL_FILL_JULIA_U));
    orc_kernel_matrix_sym(PL_KERNEL_DOT, X, n, d, d, 2, 0, NULL, 1.0, 1.0, Kref, n, 0, n);
    for (int64_t j = 0; j < n; ++j)
      for (int64_t i = 0; i < n; ++i) { /* Julia K[i+1, j+1] = K[i + j*n] */
        if (i <= j) expect_close("KernelMatrix(F, DotProduct()) upper triangle", K[i + j * n], Kref[i + j * n], i, j);
        else if (K[i + j * n] != 0.0) {
          fprintf(stderr, "FAILED: lower triangle of the Symmetric(:U) storage was written at (%ld, %ld)\n", (long)i, (long)j);
          return 1;
        }
      }
    ++checks;
    memset(K, 0, sizeof(double) * n * n);
    memset(Kref, 0, sizeof(double) * n * n);
    CALL(pl_gram_rbf(X, n, d, d, 0, NULL, 1.5, 0.7, K, n, PL_FILL_JULIA_U));
    orc_kernel_matrix_sym(PL_KERNEL_RBF, X, n, d, d, 0, 0, NULL, 1.5, 0.7, Kref, n, 0, n);
    for (int64_t j = 0; j < n; ++j)
      for (int64_t i = 0; i <= j; ++i) expect_close("KernelMatrix(F, RBF(Euclidean(d)))", K[i + j * n], Kref[i + j * n], i, j);
    for (int64_t i = 0; i < n; ++i)
      if (K[i + i * n] != 0.7) {
        fprintf(stderr, "FAILED: RBF diagonal is not beta exactly\n");
        return 1;
      }
    ++checks;
    /* KernelMatrix(F1, F2, k) -> Matrix m x n column-major == row-major n x m: the shim passes F2 first */
    const int64_t m = 57, n2 = 140;
    double* Kx = (double*)calloc(m * n2, sizeof(double));
    double* Kxr = (double*)calloc(m * n2, sizeof(double));
    const double *F1 = X, *F2 = X + m * d;
    CALL(pl_gram_cross(PL_KERNEL_RBF, F2, n2, d, F1, m, d, d, 0, 0, NULL, 1.5, 0.7, Kx, m));
    orc_kernel_matrix_cross(PL_KERNEL_RBF, F1, m, d, F2, n2, d, d, 0, 0, NULL, 1.5, 0.7, Kxr, m);
    for (int64_t j = 0; j < n2; ++j)
      for (int64_t i = 0; i < m; ++i) expect_close("KernelMatrix(F1, F2, k)", Kx[i + j * m], Kxr[i + j * m], i, j);
    ++checks;
    free(X); free(mu); free(K); free(Kref); free(Kx); free(Kxr);
  }

  /* ---- learn!(lp::CovariateLinearProblem, ws, int; λ): A = [B ; dB] as K x R column-major == row-major R x K ---- */
  {
    const int64_t ncfg = 40, natoms = 9, Kf = 37, ne = ncfg, R = ncfg * (1 + 3 * natoms);
    double* A = (double*)malloc(sizeof(double) * R * Kf);
    double* b = (double*)malloc(sizeof(double) * R);
    for (int64_t e = 0; e < R * Kf; ++e) A[e] = nrand();
    for (int64_t r = 0; r < R; ++r) b[r] = nrand();
    const double we = 100.0, wf = 1.0, lambda = 0.01;
    const int64_t n = Kf + 1;
    double* M = (double*)calloc(n * n, sizeof(double));
    double* v = (double*)calloc(n, sizeof(double));
    CALL(pl_normal_eq(A, R, Kf, Kf, NULL, ne, we, wf, b, 1, lambda, M, v));
    /* reference: A = hcat(int_col, A) with int_col = [ones(N); zeros(M)] (linear-learn.jl:239-241), Q = Diagonal (:246-249), (A'QA + λI, A'Qb) (:253) */
    double* Ai = (double*)malloc(sizeof(double) * R * n);
    double* q = (double*)malloc(sizeof(double) * R);
    for (int64_t r = 0; r < R; ++r) {
      Ai[r * n] = r < ne ? 1.0 : 0.0;
      memcpy(Ai + r * n + 1, A + r * Kf, sizeof(double) * Kf);
      q[r] = r < ne ? we : wf;
    }
    double* Mr = (double*)calloc(n * n, sizeof(double));
    double* vr = (double*)calloc(n, sizeof(double));
    orc_atqa(Ai, R, n, n, q, b, Mr, vr);
    for (int64_t i = 0; i < n; ++i) Mr[i * n + i] += lambda;
    expect_norm_close("learn! A'QA + lambda I", M, Mr, n * n);
    expect_norm_close("learn! A'Qb", v, vr, n);
    for (int64_t i = 0; i < n; ++i)
      for (int64_t j = 0; j < i; ++j)
        if (M[i * n + j] != M[j * n + i]) {
          fprintf(stderr, "FAILED: AtWA is not exactly symmetric\n");
          return 1;
        }
    /* the blocks form: pointers to lp.B packed (K x N) and to every lp.dB[i] (K x 3natoms), as they lie in Julia memory */
    const double** blocks = (const double**)malloc(sizeof(double*) * (1 + ncfg));
    int64_t* rows = (int64_t*)malloc(sizeof(int64_t) * (1 + ncfg));
    blocks[0] = A;
    rows[0] = ne;
    for (int64_t c = 0; c < ncfg; ++c) {
      blocks[1 + c] = A + (ne + c * 3 * natoms) * Kf;
      rows[1 + c] = 3 * natoms;
    }
    memset(M, 0, sizeof(double) * n * n);
    memset(v, 0, sizeof(double) * n);
    CALL(pl_normal_eq_blocks(blocks, rows, 1 + ncfg, Kf, NULL, ne, we, wf, b, 1, lambda, 0, M, v));
    expect_norm_close("learn! (blocks) A'QA + lambda I", M, Mr, n * n);
    expect_norm_close("learn! (blocks) A'Qb", v, vr, n);
    free(A); free(b); free(M); free(v); free(Ai); free(q); free(Mr); free(vr); free(blocks); free(rows);
  }

  /* ---- fit!(ds, ::PCAState): d = reduce(hcat, rows) (K x n), pca.m empty -> mean returned, centred covariance ---- */
  {
    const int64_t n = 500, Kf = 31;
    double* X = (double*)malloc(sizeof(double) * n * Kf);
    double* shift = (double*)malloc(sizeof(double) * Kf);
    for (int64_t k = 0; k < Kf; ++k) shift[k] = 10.0 * nrand();
    for (int64_t i = 0; i < n; ++i)
      for (int64_t k = 0; k < Kf; ++k) X[i * Kf + k] = shift[k] + nrand();
    double* C = (double*)calloc(Kf * Kf, sizeof(double));
    double* m = (double*)calloc(Kf, sizeof(double));
    CALL(pl_cov(X, n, Kf, Kf, m, 0, 1, C));
    double* mr = (double*)calloc(Kf, sizeof(double));
    for (int64_t i = 0; i < n; ++i) /* sum(d) / length(d) (pca_state.jl:35) */
      for (int64_t k = 0; k < Kf; ++k) mr[k] += X[i * Kf + k];
    for (int64_t k = 0; k < Kf; ++k) mr[k] /= (double)n;
    double* Cr = (double*)calloc(Kf * Kf, sizeof(double));
    orc_outer_sum(X, n, Kf, Kf, NULL, NULL, mr, Cr, NULL); /* mean(di*di') numerator, dimension_reduction.jl:12 */
    for (int64_t e = 0; e < Kf * Kf; ++e) Cr[e] /= (double)n;
    expect_norm_close("fit!(::PCAState) mean", m, mr, Kf);
    expect_norm_close("fit!(::PCAState) covariance", C, Cr, Kf * Kf);
    free(X); free(shift); free(C); free(m); free(mr); free(Cr);
  }

  pl_finalize();
  printf("OK %d\n", checks);
  return 0;
}
