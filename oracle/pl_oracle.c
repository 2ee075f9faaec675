/*
 * pl_oracle.c — CPU restatement (plain C, strict IEEE fp64, single thread) of the reference algorithms on the
 * hot path of PotentialLearning.jl v0.2.7. TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this; the product (libplb200.so, plb200/) never does.
 *
 * PARITY UNPINNED: the reference is pure Julia and Julia is not installed in the build image nor on the GPU box,
 * and the reference's own tests hold no golden values for this path. These functions follow
 * the reference line by line in the reference's own operation order (citations are into /root/reference), and
 * are cross-checked against an independent extended-precision evaluation (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -fPIC -shared -fno-fast-math -ffp-contract=off (oracle/Makefile). -ffp-contract=off keeps a*b+c
 * as two roundings like Julia's generic code; BLAS-backed reductions in the reference (dot, gemm) have unspecified
 * summation order, so agreement with the real reference is to a few ulp, not bitwise.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* dot(A,B) — LinearAlgebra.dot on Vector{Float64} (BLAS ddot in the reference; sequential here) */
static double dot(const double* a, const double* b, int64_t d) {
  double s = 0.0;
  for (int64_t k = 0; k < d; ++k) s += a[k] * b[k];
  return s;
}

/* x^p for Int p: Julia's Base.power_by_squaring semantics for p >= 0 (src/Kernels/kernels.jl:35 `^d.α`) */
static double powi(double x, int p) {
  if (p < 0) return 1.0 / powi(x, -p);
  double r = 1.0;
  while (p) {
    if (p & 1) r *= x;
    x *= x;
    p >>= 1;
  }
  return r;
}

/* compute_kernel(A,B,::DotProduct) src/Kernels/kernels.jl:30-36:  (dot(A,B)/sqrt(dot(A,A)*dot(B,B)))^alpha */
double orc_kernel_dot(const double* a, const double* b, int64_t d, int alpha) {
  return powi(dot(a, b, d) / sqrt(dot(a, a, d) * dot(b, b, d)), alpha);
}

/* compute_distance(B1,B2,::Euclidean) src/Kernels/distances.jl:58-60:  (B1-B2)' * Cinv * (B1-B2)
 * cinv_kind 0: Cinv = 1.0*I(dim) (distances.jl:43-45); 1: Diagonal (d values); 2: Symmetric d x d.
 * tmp: d doubles of scratch for the difference vector (the reference allocates it per call). */
double orc_distance_euclidean(const double* b1, const double* b2, int64_t d, int cinv_kind, const double* cinv, double* tmp) {
  for (int64_t k = 0; k < d; ++k) tmp[k] = b1[k] - b2[k];
  double s = 0.0;
  if (cinv_kind == 0) {
    for (int64_t k = 0; k < d; ++k) s += tmp[k] * (1.0 * tmp[k]);
  } else if (cinv_kind == 1) {
    for (int64_t k = 0; k < d; ++k) s += tmp[k] * (cinv[k] * tmp[k]);
  } else {
    /* (v' * C) * v evaluated left to right as Julia does for Adjoint*Matrix*Vector: row vector times matrix first */
    for (int64_t j = 0; j < d; ++j) {
      double rj = 0.0;
      for (int64_t k = 0; k < d; ++k) rj += tmp[k] * cinv[k * d + j];
      s += rj * tmp[j];
    }
  }
  return s;
}

/* compute_kernel(A,B,::RBF) src/Kernels/kernels.jl:68-75:  beta * exp(-0.5 * d2 / ell^2); RBF.alpha is never used */
double orc_kernel_rbf(const double* a, const double* b, int64_t d, int cinv_kind, const double* cinv, double ell, double beta, double* tmp) {
  const double d2 = orc_distance_euclidean(a, b, d, cinv_kind, cinv, tmp);
  return beta * exp(-0.5 * d2 / (ell * ell));
}

/* compute_kernel(A,B,::InverseMultiquadric) src/Kernels/kernels.jl:156-164:  (c2 + d2 / ell^2)^(-0.5) */
double orc_kernel_imq(const double* a, const double* b, int64_t d, int cinv_kind, const double* cinv, double ell, double c2, double* tmp) {
  const double d2 = orc_distance_euclidean(a, b, d, cinv_kind, cinv, tmp);
  return pow(c2 + d2 / (ell * ell), -0.5);
}

static double eval_kernel(int kernel_id, const double* a, const double* b, int64_t d, int alpha, int cinv_kind, const double* cinv, double ell,
                          double beta, double* tmp) {
  if (kernel_id == 1) return orc_kernel_dot(a, b, d, alpha);
  if (kernel_id == 3) return orc_kernel_imq(a, b, d, cinv_kind, cinv, ell, beta, tmp); /* beta carries c2 */
  return orc_kernel_rbf(a, b, d, cinv_kind, cinv, ell, beta, tmp);
}

/* KernelMatrix(F,k) src/Kernels/kernels.jl:211-223: K = zeros(n,n); for i=1:n, j=i:n K[i,j] = k(F_i,F_j); Symmetric(K).
 * X: row-major n x d. K: n x n, Julia column-major K[i + j*ldk] (== row-major element [j][i]); only j >= i is written,
 * the other triangle is left as the caller initialised it (zeros in the reference). kernel_id 1 = DotProduct, 2 = RBF, 3 = InverseMultiquadric (beta carries c2).
 * Rows [row_begin, row_end) only: lets the benchmark time a bounded sample of the pair loop. */
int orc_kernel_matrix_sym(int kernel_id, const double* X, int64_t n, int64_t d, int64_t ldx, int alpha, int cinv_kind, const double* cinv,
                          double ell, double beta, double* K, int64_t ldk, int64_t row_begin, int64_t row_end) {
  double* tmp = (double*)malloc(sizeof(double) * (size_t)(d > 0 ? d : 1));
  if (!tmp) return 1;
  for (int64_t i = row_begin; i < row_end; ++i) {
    for (int64_t j = i; j < n; ++j) {
      const double* a = X + i * ldx;
      const double* b = X + j * ldx;
      K[i + j * ldk] = eval_kernel(kernel_id, a, b, d, alpha, cinv_kind, cinv, ell, beta, tmp);
    }
  }
  free(tmp);
  return 0;
}

/* KernelMatrix(F1,F2,k) src/Kernels/kernels.jl:229-244: full m x n loop. K column-major m x n: K[i + j*ldk]. */
int orc_kernel_matrix_cross(int kernel_id, const double* X1, int64_t m, int64_t ldx1, const double* X2, int64_t n, int64_t ldx2, int64_t d,
                            int alpha, int cinv_kind, const double* cinv, double ell, double beta, double* K, int64_t ldk) {
  double* tmp = (double*)malloc(sizeof(double) * (size_t)(d > 0 ? d : 1));
  if (!tmp) return 1;
  for (int64_t i = 0; i < m; ++i)
    for (int64_t j = 0; j < n; ++j) {
      const double* a = X1 + i * ldx1;
      const double* b = X2 + j * ldx2;
      K[i + j * ldk] = eval_kernel(kernel_id, a, b, d, alpha, cinv_kind, cinv, ell, beta, tmp);
    }
  free(tmp);
  return 0;
}

/* sum(v*v' for v in rows) and sum(v*b ...) — src/Learning/linear-learn.jl:16-17 (and :45-49, :94-95, ...), and the
 * numerator of mean(di*di' for di in d) src/DimensionReduction/dimension_reduction.jl:12: one K x K outer product per
 * row, accumulated in row order. w == NULL -> weight 1; mean == NULL -> no centring (rows are used as given).
 * S: Kf x Kf (symmetric, both triangles), v: Kf or NULL. */
int orc_outer_sum(const double* A, int64_t R, int64_t Kf, int64_t lda, const double* w, const double* b, const double* mean, double* S,
                  double* v) {
  double* row = (double*)malloc(sizeof(double) * (size_t)Kf);
  if (!row) return 1;
  memset(S, 0, sizeof(double) * (size_t)(Kf * Kf));
  if (v) memset(v, 0, sizeof(double) * (size_t)Kf);
  for (int64_t r = 0; r < R; ++r) {
    for (int64_t k = 0; k < Kf; ++k) row[k] = mean ? A[r * lda + k] - mean[k] : A[r * lda + k];
    const double wr = w ? w[r] : 1.0;
    for (int64_t i = 0; i < Kf; ++i) {
      const double wi = wr * row[i];
      for (int64_t j = 0; j < Kf; ++j) S[i * Kf + j] += wi * row[j];
    }
    if (v && b)
      for (int64_t i = 0; i < Kf; ++i) v[i] += (wr * row[i]) * b[r];
  }
  free(row);
  return 0;
}

/* A'*Q*A and A'*Q*b as the WLS path forms them, src/Learning/linear-learn.jl:246-253: Q = Diagonal(q) is applied
 * to A first (Q*A materialised), then A' * (Q*A) as one gemm (k-loop innermost = dot-product order).
 * A: row-major R x n (n includes the intercept column if the caller added one). */
int orc_atqa(const double* A, int64_t R, int64_t n, int64_t lda, const double* q, const double* b, double* AtQA, double* AtQb) {
  double* QA = (double*)malloc(sizeof(double) * (size_t)(R > 0 ? R : 1) * (size_t)n);
  if (!QA) return 1;
  for (int64_t r = 0; r < R; ++r)
    for (int64_t k = 0; k < n; ++k) QA[r * n + k] = q[r] * A[r * lda + k];
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = 0; j < n; ++j) {
      double s = 0.0;
      for (int64_t r = 0; r < R; ++r) s += A[r * lda + i] * QA[r * n + j];
      AtQA[i * n + j] = s;
    }
  if (AtQb && b)
    for (int64_t i = 0; i < n; ++i) {
      double s = 0.0;
      for (int64_t r = 0; r < R; ++r) s += A[r * lda + i] * (q[r] * b[r]);
      AtQb[i] = s;
    }
  free(QA);
  return 0;
}

/* same contraction, cache-blocked over rows so that the benchmark's CPU leg is not artificially slow; identical
 * arithmetic per (i,j) up to summation order (partial sums per 256-row block). */
int orc_atqa_blocked(const double* A, int64_t R, int64_t n, int64_t lda, const double* q, double* AtQA) {
  const int64_t RB = 256;
  double* QA = (double*)malloc(sizeof(double) * (size_t)RB * (size_t)n);
  if (!QA) return 1;
  memset(AtQA, 0, sizeof(double) * (size_t)(n * n));
  for (int64_t r0 = 0; r0 < R; r0 += RB) {
    const int64_t rb = r0 + RB < R ? RB : R - r0;
    for (int64_t r = 0; r < rb; ++r)
      for (int64_t k = 0; k < n; ++k) QA[r * n + k] = q[r0 + r] * A[(r0 + r) * lda + k];
    for (int64_t r = 0; r < rb; ++r) {
      const double* ar = A + (r0 + r) * lda;
      const double* qr = QA + r * n;
      for (int64_t i = 0; i < n; ++i) {
        const double ai = ar[i];
        double* out = AtQA + i * n;
        for (int64_t j = 0; j < n; ++j) out[j] += ai * qr[j];
      }
    }
  }
  free(QA);
  return 0;
}
