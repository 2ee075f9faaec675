// prelude.cuh — HBM-bound preludes of the contractions: column sums / means, row norms, centring and the
// Cinv transform of Euclidean. Coalesced 128-bit loads; row reductions finish with warp shuffles; column
// reductions are two-stage with a fixed summation order (bitwise reproducible run to run).
//
// Reference sites: dot(A,A), dot(B,B) recomputed per pair (src/Kernels/kernels.jl:35); (B1-B2)'*Cinv*(B1-B2)
// (src/Kernels/distances.jl:58-60); pca.m = sum(d)/length(d) (src/DimensionReduction/pca_state.jl:35).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace plb {

constexpr int COLSUM_MAX_ROWS_PER_BLOCK = 512;

// partial[blockIdx.y][col] = sum of rows [blockIdx.y*rpb, ...) of column col (minus shift[col] per element if shift != NULL: the
// centred column sums rho = sum_r (a_r - m) that close the SYRK's one-sided centring — differences first, so nothing cancels).
// blockDim.x threads x 2 columns each. rpb (rows per block) is chosen by the host so that small matrices still fill the machine.
__global__ void colsum_partial_kernel(const double* __restrict__ A, int64_t R, int64_t Kf, int64_t lda, int rpb, double* __restrict__ partial,
                                      const double* __restrict__ shift) {
  const int64_t c0 = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  if (c0 >= Kf) return;
  const int64_t r0 = (int64_t)blockIdx.y * rpb;
  int64_t r1 = r0 + rpb;
  if (r1 > R) r1 = R;
  const bool pair = (c0 + 1 < Kf);
  double s0 = 0.0, s1 = 0.0;
  const double z0 = shift ? __ldg(shift + c0) : 0.0;
  const double z1 = (shift && pair) ? __ldg(shift + c0 + 1) : 0.0;
  const double* ptr = A + r0 * lda + c0;
  int64_t r = r0;
  if (pair) {
    // lda even and base 16B aligned (checked by the caller): 128-bit loads, 4 rows in flight
    for (; r + 4 <= r1; r += 4) {
      const double2 v0 = __ldg(reinterpret_cast<const double2*>(ptr));
      const double2 v1 = __ldg(reinterpret_cast<const double2*>(ptr + lda));
      const double2 v2 = __ldg(reinterpret_cast<const double2*>(ptr + 2 * lda));
      const double2 v3 = __ldg(reinterpret_cast<const double2*>(ptr + 3 * lda));
      s0 += v0.x - z0; s1 += v0.y - z1;
      s0 += v1.x - z0; s1 += v1.y - z1;
      s0 += v2.x - z0; s1 += v2.y - z1;
      s0 += v3.x - z0; s1 += v3.y - z1;
      ptr += 4 * lda;
    }
    for (; r < r1; ++r) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(ptr));
      s0 += v.x - z0; s1 += v.y - z1;
      ptr += lda;
    }
  } else {
    for (; r < r1; ++r) {
      s0 += __ldg(ptr) - z0;
      ptr += lda;
    }
  }
  partial[(int64_t)blockIdx.y * Kf + c0] = s0;
  if (pair) partial[(int64_t)blockIdx.y * Kf + c0 + 1] = s1;
}

// out[col] = (sum_b partial[b][col]) (/ denom if denom != 0). Block = 32 columns x 8 row lanes: lane rl sums partial rows
// rl, rl+8, ... (4 independent chains), then the 8 lane sums are added in fixed order: deterministic, and the pass stays
// a few microseconds even with hundreds of partial rows (a one-thread-per-column loop took 181 us at config 2).
__global__ void colsum_final_kernel(const double* __restrict__ partial, int64_t nblocks, int64_t Kf, double denom, double* __restrict__ out) {
  __shared__ double sh[8][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + cl;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (c < Kf) {
    int64_t b = rl;
    for (; b + 24 < nblocks; b += 32) {
      s0 += partial[b * Kf + c];
      s1 += partial[(b + 8) * Kf + c];
      s2 += partial[(b + 16) * Kf + c];
      s3 += partial[(b + 24) * Kf + c];
    }
    for (; b < nblocks; b += 8) s0 += partial[b * Kf + c];
  }
  sh[rl][cl] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (rl == 0 && c < Kf) {
    double tot = ((sh[0][cl] + sh[1][cl]) + (sh[2][cl] + sh[3][cl])) + ((sh[4][cl] + sh[5][cl]) + (sh[6][cl] + sh[7][cl]));
    if (denom != 0.0) tot = tot / denom;
    out[c] = tot;
  }
}

// out[col] = mean over the ns sample rows 0, stride, 2 stride, ... of column col: the SHIFT of the distance kernels (any point near the
// centre of the data removes the cancellation of the Gram form against |mean|^2; distances are translation invariant). ONE launch:
// block = 32 columns x 32 row lanes, lane rl sums sample rows rl, rl + 32, ... with four loads in flight (the pass is latency bound: every
// sample row is its own DRAM page), the 32 lane sums are added in fixed order — deterministic, the same on every rank. ~5 us for
// ns <= 512 (the two-launch colsum pass over 4096 rows took 17 us of the 530 us an 8-GPU shard of config 2 takes).
// a NaN / Inf entry of a sampled row must not poison the SHIFT (and through it every row of the result): it counts as 0 here and
// propagates through its own row only, as in the reference
__device__ __forceinline__ double fin(double v) { return (__double2hiint(v) & 0x7ff00000) == 0x7ff00000 ? 0.0 : v; }
__global__ void __launch_bounds__(1024) sample_mean_kernel(const double* __restrict__ X, int64_t ns, int64_t stride_elems, int64_t d, double* __restrict__ out) {
  pdl_launch_dependents();  // the operand kernel that follows may start its prologue; it waits (pdl_wait) before it reads the mean
  __shared__ double sh[32][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + cl;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (c < d) {
    int64_t r = rl;
    for (; r + 96 < ns; r += 128) {
      const double v0 = fin(__ldg(X + r * stride_elems + c)), v1 = fin(__ldg(X + (r + 32) * stride_elems + c));
      const double v2 = fin(__ldg(X + (r + 64) * stride_elems + c)), v3 = fin(__ldg(X + (r + 96) * stride_elems + c));
      s0 += v0; s1 += v1; s2 += v2; s3 += v3;
    }
    for (; r < ns; r += 32) s0 += fin(__ldg(X + r * stride_elems + c));
  }
  sh[rl][cl] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (rl == 0 && c < d) {
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k) tot += sh[k][cl];
    out[c] = tot / (double)ns;
  }
}

enum { PREP_DOT = 0, PREP_DIST_ID = 1, PREP_DIST_DIAG = 2, PREP_CENTER_ONLY = 3 };

// One warp per row.
//   PREP_DOT:         stat[i] = 1/sqrt(sum_k x^2)                                  (kernels.jl:35)
//   PREP_DIST_ID:     xc = x - mean -> Xc ; stat[i] = sum_k xc^2                   (distances.jl:59, Cinv = I; InverseMultiquadric — RBF has rbf_aug_kernel)
//   PREP_DIST_DIAG:   xc = x - mean -> Xc ; y = xc * cinv -> Y ; stat[i] = sum_k y*xc
//   PREP_CENTER_ONLY: xc = x - mean -> Xc                                          (then Y = Xc*Cinv by GEMM)
// ldc is the leading dimension of Xc and Y.
template <int MODE>
__global__ void prep_rows_kernel(const double* __restrict__ X, int64_t n, int64_t d, int64_t ldx, const double* __restrict__ mean,
                                 const double* __restrict__ cdiag, double* __restrict__ Xc, double* __restrict__ Y, int64_t ldc,
                                 double* __restrict__ stat) {
  pdl_launch_dependents();  // the Gram kernel that follows may start its prologue (and, for DotProduct, its main loop) now
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const double* x = X + row * ldx;
  double acc = 0.0;
  const int64_t dpair = d & ~(int64_t)1;
  for (int64_t k = 2 * lane; k < dpair; k += 64) {
    double2 v = __ldg(reinterpret_cast<const double2*>(x + k));
    if (MODE == PREP_DOT) {
      acc = fma(v.x, v.x, acc);
      acc = fma(v.y, v.y, acc);
    } else {
      const double2 m = __ldg(reinterpret_cast<const double2*>(mean + k));
      v.x -= m.x;
      v.y -= m.y;
      *reinterpret_cast<double2*>(Xc + row * ldc + k) = v;
      if (MODE == PREP_DIST_ID) {
        acc = fma(v.x, v.x, acc);
        acc = fma(v.y, v.y, acc);
      } else if (MODE == PREP_DIST_DIAG) {
        const double2 c = __ldg(reinterpret_cast<const double2*>(cdiag + k));
        double2 y;
        y.x = v.x * c.x;
        y.y = v.y * c.y;
        *reinterpret_cast<double2*>(Y + row * ldc + k) = y;
        acc = fma(y.x, v.x, acc);
        acc = fma(y.y, v.y, acc);
      }
    }
  }
  if ((d & 1) && lane == 0) {
    const int64_t k = d - 1;
    double v = __ldg(x + k);
    if (MODE == PREP_DOT) {
      acc = fma(v, v, acc);
    } else {
      v -= mean[k];
      Xc[row * ldc + k] = v;
      if (MODE == PREP_DIST_ID) acc = fma(v, v, acc);
      else if (MODE == PREP_DIST_DIAG) {
        const double y = v * cdiag[k];
        Y[row * ldc + k] = y;
        acc = fma(y, v, acc);
      }
    }
  }
  if (MODE != PREP_CENTER_ONLY) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) stat[row] = (MODE == PREP_DOT) ? 1.0 / sqrt(acc) : acc;
  }
}

// ---- RBF operands in the AUGMENTED layout (consumed by gram_kernel<EPI_RBF>, gram.cuh exp2_tab) ----
// The RBF epilogue wants y_ij = L (-0.5 d2_ij / ell^2 + ln beta), L = 256 log2(e), straight out of the accumulator. With
//   x~ = c (x - mean), c = sqrt(L) / |ell|,   q~_i = x~_i' Cinv x~_i,   u_i = -q~_i / 2 + h,  h = L ln(beta) / 2
// that is  y_ij = (Cinv x~_i) . x~_j + u_i * 1 + 1 * u_j : a contraction over d + 2 columns. One warp per row writes
//   cols [0, d)               the row (x~ for Xc, y~ = Cinv x~ for Y)
//   cols [d, 16 KT - 2)       0                                   KT = ceil((d + 2) / 16) k-steps, ld = 16 KT + 16
//   cols 16 KT - 2, 16 KT - 1 (u, 1)                              what the A operand contracts in its last k-step
//   cols [16 KT, 16 KT + 16)  the last 16-column block again, ending in (1, u): what the B operand reads as ITS last k-step
//                             (GramParams::b_shift = 16), so one buffer serves as both operands of a symmetric matrix.
// q~ is summed over the ROUNDED x~ so that identical rows give d2 = 0 up to the accumulation order. (kernels.jl:73-74, distances.jl:58-60)
enum { AUG_ID = 0, AUG_DIAG = 1, AUG_CENTER = 2, AUG_FINISH = 3 };

// tail of one row buffer: zero padding, (u, 1), and the shifted block with (1, u). `row` points at the row (ld = 16 KT + 16 doubles);
// lanes 0..15 own the columns of the last block, lane 16 the (at most one) padding column in front of it.
__device__ __forceinline__ void rbf_aug_tail(double* __restrict__ row, const int64_t d, const int KT, const double u, const int lane) {
  const int64_t b0 = 16 * (int64_t)(KT - 1);
  if (lane < 16) {
    const int64_t col = b0 + lane;
    const double v = col < d ? row[col] : 0.0;  // written by this warp before (__syncwarp in the caller)
    if (col >= d) row[col] = lane == 14 ? u : (lane == 15 ? 1.0 : 0.0);
    row[b0 + 16 + lane] = lane == 14 ? 1.0 : (lane == 15 ? u : v);
  } else if (lane == 16 && d < b0) {
    row[d] = 0.0;  // d == 16 (KT - 1) - 1: the only padding column outside the last block
  }
}

template <int MODE>
__global__ void rbf_aug_kernel(const double* __restrict__ X, int64_t n, int64_t d, int64_t ldx, const double* __restrict__ mean,
                               const double* __restrict__ cdiag, double* __restrict__ Xc, double* __restrict__ Y, int KT, double c, double h) {
  pdl_launch_dependents();  // the Gram kernel that follows may start its prologue now
  pdl_wait();               // launched as a programmatic dependent of the kernel that writes `mean` (no-op otherwise)
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const int64_t ld = 16 * (int64_t)KT + 16;
  double* const xc = Xc + row * ld;
  double* const y = (MODE == AUG_DIAG || MODE == AUG_FINISH) ? Y + row * ld : nullptr;
  double acc = 0.0;
  if (MODE == AUG_FINISH) {
    for (int64_t k = lane; k < d; k += 32) acc = fma(y[k], xc[k], acc);
  } else {
    const double* x = X + row * ldx;
    const int64_t dpair = d & ~(int64_t)1;
    for (int64_t k = 2 * lane; k < dpair; k += 64) {
      double2 v = __ldg(reinterpret_cast<const double2*>(x + k));
      const double2 m = __ldg(reinterpret_cast<const double2*>(mean + k));
      v.x = (v.x - m.x) * c;
      v.y = (v.y - m.y) * c;
      *reinterpret_cast<double2*>(xc + k) = v;
      if (MODE == AUG_ID) {
        acc = fma(v.x, v.x, acc);
        acc = fma(v.y, v.y, acc);
      } else if (MODE == AUG_DIAG) {
        const double2 cd = __ldg(reinterpret_cast<const double2*>(cdiag + k));
        double2 w;
        w.x = v.x * cd.x;
        w.y = v.y * cd.y;
        *reinterpret_cast<double2*>(y + k) = w;
        acc = fma(w.x, v.x, acc);
        acc = fma(w.y, v.y, acc);
      }
    }
    if ((d & 1) && lane == 0) {
      const int64_t k = d - 1;
      const double v = (__ldg(x + k) - mean[k]) * c;
      xc[k] = v;
      if (MODE == AUG_ID) acc = fma(v, v, acc);
      else if (MODE == AUG_DIAG) {
        const double w = v * cdiag[k];
        y[k] = w;
        acc = fma(w, v, acc);
      }
    }
  }
  if constexpr (MODE != AUG_CENTER) {  // AUG_CENTER: Y = Xc Cinv by the Gram engine next, then AUG_FINISH
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    const double u = fma(-0.5, acc, h);
    __syncwarp();
    rbf_aug_tail(xc, d, KT, u, lane);
    if (MODE == AUG_DIAG || MODE == AUG_FINISH) rbf_aug_tail(y, d, KT, u, lane);
  }
}

// stat[i] = sum_k Y[i][k] * Xc[i][k]  (quadratic form x' Cinv x with Y = Xc Cinv); one warp per row
__global__ void rowdot_kernel(const double* __restrict__ Y, const double* __restrict__ Xc, int64_t n, int64_t d, int64_t ld, double* __restrict__ stat) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  double acc = 0.0;
  for (int64_t k = lane; k < d; k += 32) acc = fma(Y[row * ld + k], Xc[row * ld + k], acc);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) stat[row] = acc;
}

__global__ void scale_div_kernel(const double* __restrict__ in, double denom, int64_t count, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i] / denom;
}

// y[r] = sum_k A[r][k] * beta[k] + (r < ne ? beta0 : 0): the prediction of a linear basis potential over the packed rows
// (energies of the first ne rows get the intercept, force rows do not — the same [1_N; 0_M] column as in the fit,
// linear-learn.jl:239-241). HBM-bound GEMV: one warp per row, 128-bit coalesced loads, beta through the read-only cache,
// warp-shuffle reduction. SURVEY.md 8f row f3 (what get_all_energies / get_all_forces(ds, lb) need after learn!).
__global__ void predict_rows_kernel(const double* __restrict__ A, int64_t R, int64_t Kf, int64_t lda, const double* __restrict__ beta,
                                    double beta0, int64_t ne, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t kpair = Kf & ~(int64_t)1;
  for (int64_t r = warp0; r < R; r += nwarps) {
    const double* a = A + r * lda;
    double acc0 = 0.0, acc1 = 0.0;
    for (int64_t k = 2 * lane; k < kpair; k += 64) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(a + k));
      const double2 b = __ldg(reinterpret_cast<const double2*>(beta + k));
      acc0 = fma(v.x, b.x, acc0);
      acc1 = fma(v.y, b.y, acc1);
    }
    if ((Kf & 1) && lane == 0) acc0 = fma(__ldg(a + Kf - 1), __ldg(beta + Kf - 1), acc0);
    double acc = acc0 + acc1;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) y[r] = acc + (r < ne ? beta0 : 0.0);
  }
}

// out[i][j] = out[j][i] = in[min(i,j)][max(i,j)] / denom: exactly symmetric K x K result (covariance 1/n after the all-reduce)
__global__ void sym_scale_div_kernel(const double* __restrict__ in, int64_t K, double denom, double* __restrict__ out) {
  const int64_t total = K * K;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / K, j = e % K;
    out[e] = (i <= j ? in[i * K + j] : in[j * K + i]) / denom;
  }
}

// ---- GlobalSum / GlobalMean features over the stacked local descriptors of all configurations ----
// X[i][:] = sum_{a in [off[i], off[i+1])} L[a][:]  (/ natoms_i for the mean): compute_feature(B, ::GlobalSum / ::GlobalMean)
// (src/Kernels/features.jl:19-21, 32-34), B[i] = sum(get_values(local descriptors)) (linear-learning-problem.jl:75) and the rows
// of fit!(::PCAState) (pca_state.jl:23). HBM-bound segmented reduction: a group of G lanes (G = 1..32, a power of two) owns one
// (configuration, 2G-column chunk); each lane carries one 16-byte column pair through the atoms with 4 loads in flight.
// The summation ORDER is Julia's: Base.sum over a Vector{Vector} is mapreduce_impl with pairwise_blocksize 1024 — a plain
// left-to-right sum for fewer than 1024 atoms, and above that a recursive split at imid = ifirst + (ilast - ifirst) >> 1 whose
// leaves are summed left to right. The same tree is walked here with an explicit stack, so the result is bit-identical to the
// reference's for every natoms (the mean divides the finished sum by natoms, as Statistics.mean does).
struct SegFrame {
  int64_t lo, hi;
  double2 v1;
  int stage;
};
__device__ __forceinline__ double2 seg_sum_seq(const double* __restrict__ base, int64_t ld, int64_t lo, int64_t hi, bool pair) {
  // rows lo..hi inclusive, sequential; `base` already points at this lane's column pair
  double2 s;
  const double* p = base + lo * ld;
  if (pair) {
    s = __ldg(reinterpret_cast<const double2*>(p));
    int64_t r = lo + 1;
    p += ld;
    for (; r + 3 <= hi; r += 4) {
      const double2 a0 = __ldg(reinterpret_cast<const double2*>(p));
      const double2 a1 = __ldg(reinterpret_cast<const double2*>(p + ld));
      const double2 a2 = __ldg(reinterpret_cast<const double2*>(p + 2 * ld));
      const double2 a3 = __ldg(reinterpret_cast<const double2*>(p + 3 * ld));
      s.x += a0.x; s.y += a0.y;
      s.x += a1.x; s.y += a1.y;
      s.x += a2.x; s.y += a2.y;
      s.x += a3.x; s.y += a3.y;
      p += 4 * ld;
    }
    for (; r <= hi; ++r) {
      const double2 a = __ldg(reinterpret_cast<const double2*>(p));
      s.x += a.x; s.y += a.y;
      p += ld;
    }
  } else {
    s.x = __ldg(p);
    s.y = 0.0;
    for (int64_t r = lo + 1; r <= hi; ++r) {
      p += ld;
      s.x += __ldg(p);
    }
  }
  return s;
}
constexpr int SEG_PAIRWISE_BLOCK = 1024;  // Base.pairwise_blocksize(identity, +)
__global__ void segment_reduce_kernel(const double* __restrict__ L, int64_t ldl, const int64_t* __restrict__ off, int64_t N, int64_t d, int G,
                                      int mean, double* __restrict__ X, int64_t ldx) {
  const int64_t gthread = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t group = gthread / G;
  const int gl = (int)(gthread % G);
  const int64_t nchunks = (d + 2 * G - 1) / (2 * G);
  const int64_t cfg = group / nchunks;
  if (cfg >= N) return;
  const int64_t c0 = (group % nchunks) * 2 * G + 2 * gl;
  if (c0 >= d) return;
  const bool pair = c0 + 1 < d;
  const int64_t lo = off[cfg], hi = off[cfg + 1] - 1;
  double2 s;
  const double* base = L + c0;
  if (hi < lo) {
    s.x = s.y = 0.0;  // empty configuration: the reference's sum/mean of an empty collection throws; the host rejects it
  } else if (hi - lo < SEG_PAIRWISE_BLOCK) {
    s = seg_sum_seq(base, ldl, lo, hi, pair);
  } else {
    SegFrame st[48];
    int sp = 0;
    st[0].lo = lo; st[0].hi = hi; st[0].stage = 0;
    double2 ret;
    ret.x = ret.y = 0.0;
    while (sp >= 0) {
      SegFrame& f = st[sp];
      if (f.stage == 0) {
        if (f.hi - f.lo < SEG_PAIRWISE_BLOCK) {
          ret = seg_sum_seq(base, ldl, f.lo, f.hi, pair);
          --sp;
        } else {
          const int64_t mid = f.lo + ((f.hi - f.lo) >> 1);
          f.stage = 1;
          st[sp + 1].lo = f.lo; st[sp + 1].hi = mid; st[sp + 1].stage = 0;
          ++sp;
        }
      } else if (f.stage == 1) {
        f.v1 = ret;
        const int64_t mid = f.lo + ((f.hi - f.lo) >> 1);
        f.stage = 2;
        st[sp + 1].lo = mid + 1; st[sp + 1].hi = f.hi; st[sp + 1].stage = 0;
        ++sp;
      } else {
        ret.x = f.v1.x + ret.x;
        ret.y = f.v1.y + ret.y;
        --sp;
      }
    }
    s = ret;
  }
  if (mean) {
    const double n = (double)(hi - lo + 1);
    s.x /= n;
    s.y /= n;
  }
  double* out = X + cfg * ldx + c0;
  if (pair && !(ldx & 1) && !(reinterpret_cast<uintptr_t>(X) & 15u)) *reinterpret_cast<double2*>(out) = s;
  else {
    out[0] = s.x;
    if (pair) out[1] = s.y;
  }
}

// acc[i] += x[i]: accumulation of the chunk partials of the streamed normal equations in fixed (chunk) order
__global__ void accumulate_kernel(const double* __restrict__ x, int64_t count, double* __restrict__ acc) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) acc[i] += x[i];
}

// ---- L-ensemble helpers on the device-resident eigenvectors ----
// V: row k = eigenvector k (n contiguous doubles, leading dimension ldv). partial[blockIdx.y][i] = sum over this block's
// eigenvector range of gamma[k] * V[k][i]^2; colsum_final_kernel adds the partials in fixed order:
// inclusion_prob(L) = diag(U diag(a*lam/(1+a*lam)) U') (Determinantal.jl [external], called at src/SubsetSelection/dpp.jl:69).
__global__ void weighted_colsumsq_partial_kernel(const double* __restrict__ V, int64_t nvec, int64_t n, int64_t ldv, const double* __restrict__ gamma,
                                                 int rpb, double* __restrict__ partial) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t k0 = (int64_t)blockIdx.y * rpb;
  int64_t k1 = k0 + rpb;
  if (k1 > nvec) k1 = nvec;
  double s0 = 0.0, s1 = 0.0;
  int64_t k = k0;
  for (; k + 2 <= k1; k += 2) {
    const double a = __ldg(V + k * ldv + i), b = __ldg(V + (k + 1) * ldv + i);
    s0 = fma(__ldg(gamma + k) * a, a, s0);
    s1 = fma(__ldg(gamma + k + 1) * b, b, s1);
  }
  if (k < k1) {
    const double a = __ldg(V + k * ldv + i);
    s0 = fma(__ldg(gamma + k) * a, a, s0);
  }
  partial[(int64_t)blockIdx.y * n + i] = s0 + s1;
}

// gamma[k] = a*max(lam[k],0) / (1 + a*max(lam[k],0))
__global__ void ell_gamma_kernel(const double* __restrict__ lam, int64_t n, double alpha, double* __restrict__ gamma) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double l = lam[k];
  l = l > 0.0 ? l : 0.0;
  const double g = alpha * l;
  gamma[k] = g / (1.0 + g);
}

// out[r][:] = V[idx[r]][:]: the eigenvectors the k-DPP sampler selected, packed for one D2H copy
__global__ void gather_rows_kernel(const double* __restrict__ V, int64_t n, int64_t ldv, const int64_t* __restrict__ idx, int64_t m, double* __restrict__ out) {
  const int64_t r = blockIdx.y;
  if (r >= m) return;
  const double* src = V + idx[r] * ldv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[r * n + i] = src[i];
}

// Weighted column sums of a row block — the moment vectors of the normal equations (an HBM-rate pass of their own: inside the DMMA
// loop of the SYRK every scalar FP64 instruction costs about a DMMA, see syrk.cuh):
//   pv[blk][col] = sum_r w_r b_r A[r][col]   (A'Wb, linear-learn.jl:200, 253)        — only if b != NULL
//   pc[blk][col] = sum_{r < ne} w_r A[r][col] (the intercept row/column [1_N; 0_M]' W A, linear-learn.jl:239-241)
// w_r = w[r] if w != NULL else (r < ne ? w_e : w_f). Same geometry as colsum_partial_kernel; colsum_final_kernel adds the blocks.
__global__ void wcolsum_partial_kernel(const double* __restrict__ A, int64_t R, int64_t Kf, int64_t lda, const double* __restrict__ w, int64_t ne,
                                       double w_e, double w_f, const double* __restrict__ b, int rpb, double* __restrict__ pv, double* __restrict__ pc) {
  const int64_t c0 = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  if (c0 >= Kf) return;
  const int64_t r0 = (int64_t)blockIdx.y * rpb;
  int64_t r1 = r0 + rpb;
  if (r1 > R) r1 = R;
  const bool pair = (c0 + 1 < Kf);
  double v0 = 0.0, v1 = 0.0, s0 = 0.0, s1 = 0.0;
  const double* ptr = A + r0 * lda + c0;
  auto row = [&](int64_t r, double2 a) {
    const double wr = w ? __ldg(w + r) : (r < ne ? w_e : w_f);
    if (b) {
      const double wb = wr * __ldg(b + r);
      v0 = fma(a.x, wb, v0);
      v1 = fma(a.y, wb, v1);
    }
    if (r < ne) {
      s0 = fma(a.x, wr, s0);
      s1 = fma(a.y, wr, s1);
    }
  };
  int64_t r = r0;
  if (pair) {
    for (; r + 4 <= r1; r += 4) {
      const double2 a0 = __ldg(reinterpret_cast<const double2*>(ptr));
      const double2 a1 = __ldg(reinterpret_cast<const double2*>(ptr + lda));
      const double2 a2 = __ldg(reinterpret_cast<const double2*>(ptr + 2 * lda));
      const double2 a3 = __ldg(reinterpret_cast<const double2*>(ptr + 3 * lda));
      row(r, a0);
      row(r + 1, a1);
      row(r + 2, a2);
      row(r + 3, a3);
      ptr += 4 * lda;
    }
    for (; r < r1; ++r) {
      row(r, __ldg(reinterpret_cast<const double2*>(ptr)));
      ptr += lda;
    }
  } else {
    for (; r < r1; ++r) {
      double2 a;
      a.x = __ldg(ptr);
      a.y = 0.0;
      row(r, a);
      ptr += lda;
    }
  }
  if (b) {
    pv[(int64_t)blockIdx.y * Kf + c0] = v0;
    if (pair) pv[(int64_t)blockIdx.y * Kf + c0 + 1] = v1;
  }
  pc[(int64_t)blockIdx.y * Kf + c0] = s0;
  if (pair) pc[(int64_t)blockIdx.y * Kf + c0 + 1] = s1;
}

// CorrelationMatrix features (src/Kernels/features.jl:48-50): C_i = mean(b b' for b in B_i), the d x d second moment of configuration
// i's local descriptors. One CTA per configuration, one thread per (a, b) entry (strided). The reference forms each outer product
// (rounded products) and adds the matrices in atom order — a generator, so Statistics.mean accumulates sequentially — then divides by
// natoms: separate multiply and add here (no fma contraction) in the same order, so the features are bit-identical.
// out: N x d x d row-major (full symmetric matrices).
__global__ void correlation_feature_kernel(const double* __restrict__ L, int64_t ldl, const int64_t* __restrict__ off, int64_t d,
                                           double* __restrict__ out) {
  const int64_t cfg = blockIdx.x;
  const int64_t lo = off[cfg], hi = off[cfg + 1];
  const double n = (double)(hi - lo);
  double* C = out + cfg * d * d;
  for (int64_t e = threadIdx.x; e < d * d; e += blockDim.x) {
    const int64_t a = e / d, b = e % d;
    double s = 0.0;
    if (hi > lo) {
      s = __dmul_rn(__ldg(L + lo * ldl + a), __ldg(L + lo * ldl + b));
      for (int64_t r = lo + 1; r < hi; ++r) s = __dadd_rn(s, __dmul_rn(__ldg(L + r * ldl + a), __ldg(L + r * ldl + b)));
    }
    C[e] = s / n;
  }
}

// ---- KernelSteinDiscrepancy with an RBF(Euclidean) kernel (src/Kernels/divergences.jl:27-49) ----
// The reference's double loop evaluates, per pair i <= j with u = x_i - x_j, C = Cinv, k = beta exp(-u'Cu / 2l^2):
//   sks = s_i.s_j k ;  sk = s_i'(grad_y k) = k s_i'Cu / l^2 ;  ks = (grad_x k)'s_j = -k u'C s_j / l^2 ;
//   trk = tr(k .* (C/l^2 .- (u'C^2 u)/l^4)) = k (tr C / l^2 - D u'C^2u / l^4)
// (kernels.jl:81-126, distances.jl:76-110; the `.-` broadcasts the scalar onto EVERY entry of the D x D matrix, so the trace picks it
// up D times — kept as written), and sums m (sks + sk + ks + trk) / (N (N-1)), m = 1 on the diagonal, 2 off it.
// With Y = Xc C (Xc = mean-shifted X: u is translation invariant) all five pair quantities are entries of four Gram matrices built by
// the DMMA engine: G1 = Xc Y', G2 = Y Y', G3 = S S', H = S Y':
//   u'Cu = G1_ii + G1_jj - 2 G1_ij,  u'C^2u = G2_ii + G2_jj - 2 G2_ij,  s_i.s_j = G3_ij,  s_i'Cu - u'Cs_j = H_ii - H_ij - H_ji + H_jj.
// This kernel is the fused pair reduction over the upper triangle: one partial sum per block (fixed order inside the block),
// added by ksd_final_kernel in block order -> deterministic.
__global__ void ksd_reduce_kernel(const double* __restrict__ G1, const double* __restrict__ G2, const double* __restrict__ G3,
                                  const double* __restrict__ H, int64_t n, int64_t ld, double beta, double inv_l2, double trC, double D,
                                  double* __restrict__ partial) {
  __shared__ double sh[256];
  const int64_t i = blockIdx.y;
  double acc = 0.0;
  const double g1ii = G1[i * ld + i], g2ii = G2[i * ld + i], hii = H[i * ld + i];
  for (int64_t j = i + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const double d2 = g1ii + G1[j * ld + j] - 2.0 * G1[i * ld + j];
    const double c2 = g2ii + G2[j * ld + j] - 2.0 * G2[i * ld + j];
    const double k = beta * exp(-0.5 * d2 * inv_l2);
    const double cross = hii - H[i * ld + j] - H[j * ld + i] + H[j * ld + j];
    const double t = k * (G3[i * ld + j] + cross * inv_l2 + trC * inv_l2 - D * c2 * inv_l2 * inv_l2);
    acc += (j == i) ? t : 2.0 * t;
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = sh[0];
}
__global__ void ksd_final_kernel(const double* __restrict__ partial, int64_t count, double denom, double* __restrict__ out) {
  __shared__ double sh[1024];
  double a = 0.0;
  for (int64_t k = threadIdx.x; k < count; k += blockDim.x) a += partial[k];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0] / denom;
}
// Y = X .* c (row-wise scaling by a diagonal Cinv) for an already centred X
__global__ void scale_cols_kernel(const double* __restrict__ X, int64_t n, int64_t d, int64_t ld, const double* __restrict__ c, double* __restrict__ Y) {
  const int64_t total = n * d;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / d, k = e % d;
    Y[r * ld + k] = X[r * ld + k] * c[k];
  }
}

// S_ij <- S_ij - rho_i delta_j - delta_i rho_j + n delta_i delta_j over the j >= i triangle, mirrored: re-centres a second moment taken
// about a shift z (S = sum (x-z)(x-z)', rho = sum (x-z)) to the mean m = z + delta. Exactly symmetric by construction.
__global__ void cov_recentre_kernel(double* __restrict__ S, int64_t K, const double* __restrict__ rho, const double* __restrict__ delta, double n) {
  const int64_t total = K * K;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / K, j = e % K;
    if (j < i) continue;
    double v = S[e];
    v = fma(-rho[i], delta[j], v);
    v = fma(-delta[i], rho[j], v);
    v = fma(n * delta[i], delta[j], v);
    S[e] = v;
    S[j * K + i] = v;
  }
}
// out[i] = a[i] + b[i]
__global__ void axpy1_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t count, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = a[i] + b[i];
}

// ---- first phase of a k-DPP draw: which eigenvectors (Kulesza & Taskar 2012, Alg. 7 + Alg. 8), in the log domain ----
// T[n][l] = log e_l(lam_1..lam_n): T[n][l] = logaddexp(T[n-1][l], log lam_n + T[n-1][l-1]). Row n depends on row n-1 only, so one
// thread-block CLUSTER (8 CTAs x 1024 threads) walks the rows with a cluster barrier per row; the l range is spread over the cluster.
// Rows are exchanged through global memory (L2 loads: another SM's L1 is not coherent); the table is kept for the selection pass.
constexpr int ESP_CLUSTER = 8;
__global__ void esp_loglam_kernel(const double* __restrict__ lam, int64_t N, double* __restrict__ loglam) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) loglam[i] = lam[i] > 0.0 ? log(lam[i]) : -INFINITY;
}
__global__ void __cluster_dims__(ESP_CLUSTER, 1, 1) __launch_bounds__(1024) esp_forward_kernel(const double* __restrict__ loglam, int64_t N, int64_t k, double* T, int64_t ldt) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int64_t gtid = (int64_t)cluster.block_rank() * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)ESP_CLUSTER * blockDim.x;
  for (int64_t l = gtid; l <= k; l += nth) T[l] = l == 0 ? 0.0 : -INFINITY;
  cluster.sync();
  for (int64_t n = 1; n <= N; ++n) {
    const double ll = loglam[n - 1];
    const double* prev = T + (n - 1) * ldt;
    double* cur = T + n * ldt;
    for (int64_t l = gtid; l <= k; l += nth) {
      double v;
      if (l == 0) {
        v = 0.0;
      } else if (l > n) {
        v = -INFINITY;  // e_l of fewer than l values
      } else {
        const double a = __ldcg(prev + l), b = ll + __ldcg(prev + l - 1);
        const double mx = fmax(a, b);
        v = (mx == -INFINITY) ? mx : mx + log1p(exp(-fabs(a - b)));
      }
      cur[l] = v;
    }
    cluster.sync();
  }
}
// Alg. 8, backwards over the eigenvalues with one uniform per visited eigenvalue (uniforms[N - n] for eigenvalue n): eigenvector n-1 is
// taken with probability lam_n e_{l-1}^{n-1} / e_l^n. idx: the taken eigenvectors in visiting order (descending); count[0] = how many.
__global__ void esp_select_kernel(const double* __restrict__ T, int64_t ldt, const double* __restrict__ loglam, int64_t N, int64_t k,
                                  const double* __restrict__ uniforms, int64_t* __restrict__ idx, int64_t* __restrict__ count) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int64_t l = k, cnt = 0;
  for (int64_t n = N; n >= 1 && l > 0; --n) {
    if (l == n) {  // every remaining eigenvector is needed
      for (int64_t r = n - 1; r >= 0; --r) idx[cnt++] = r;
      l = 0;
      break;
    }
    const double logp = loglam[n - 1] + T[(n - 1) * ldt + l - 1] - T[n * ldt + l];
    if (log(uniforms[N - n]) < logp) {
      idx[cnt++] = n - 1;
      --l;
    }
  }
  count[0] = cnt;
}

// ---- projection-DPP sampling over the selected eigenvectors of the resident L-ensemble (second phase of a k-DPP draw) ----
// Kulesza & Taskar 2012, Alg. 1 in its O(n m^2) Gram-Schmidt form (what Determinantal.jl's sample_pdpp does): with V the n x m matrix of
// the selected eigenvectors (stored transposed: Vt is m x n, row k = eigenvector k) and p_j = |V_j|^2,
//   repeat m times:  item ~ p / sum(p);  v = V[item, :];  f = v - F F'v (F = the orthonormal directions chosen so far);
//                    f /= |f|;  p_j -= (V_j . f)^2.
// The uniforms come from the caller (the RNG stays in Julia / numpy); every step is a handful of small kernels plus one pass over Vt,
// enqueued without host synchronisation (the picked item lives in device memory).
__global__ void dpp_init_p_kernel(const double* __restrict__ Vt, int64_t m, int64_t n, double* __restrict__ p) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double s0 = 0.0, s1 = 0.0;
  int64_t k = 0;
  for (; k + 2 <= m; k += 2) {
    const double a = Vt[k * n + j], b = Vt[(k + 1) * n + j];
    s0 = fma(a, a, s0);
    s1 = fma(b, b, s1);
  }
  if (k < m) s0 = fma(Vt[k * n + j], Vt[k * n + j], s0);
  p[j] = s0 + s1;
}

// one block: item = the first j with cumsum(p)[j] > u * sum(p) (inverse CDF; deterministic chunked scan)
__global__ void dpp_pick_kernel(const double* __restrict__ p, int64_t n, const double* __restrict__ uniforms, int64_t step, int64_t* __restrict__ items) {
  __shared__ double part[1024];
  __shared__ double total_sh;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t chunk = (n + nt - 1) / nt;
  const int64_t j0 = (int64_t)tid * chunk, j1 = j0 + chunk < n ? j0 + chunk : n;
  double s = 0.0;
  for (int64_t j = j0; j < j1; ++j) s += p[j] > 0.0 ? p[j] : 0.0;
  part[tid] = s;
  __syncthreads();
  if (tid == 0) {  // exclusive prefix over the (<= 1024) chunk sums: sequential, negligible next to the pass over Vt
    double run = 0.0;
    for (int k = 0; k < nt; ++k) {
      const double x = part[k];
      part[k] = run;
      run += x;
    }
    total_sh = run;
  }
  __syncthreads();
  const double target = uniforms[step] * total_sh;
  const double lo = part[tid], hi = lo + s;
  const bool last_nonempty = (hi >= total_sh) && (s > 0.0);  // guards u -> 1 against rounding
  if ((target >= lo && target < hi) || (last_nonempty && target >= hi)) {
    double run = lo;
    int64_t pick = -1;
    for (int64_t j = j0; j < j1; ++j) {
      const double x = p[j] > 0.0 ? p[j] : 0.0;
      if (x > 0.0) pick = j;  // remember the last admissible item of the chunk
      run += x;
      if (run > target && x > 0.0) {
        pick = j;
        break;
      }
    }
    if (pick >= 0) items[step] = pick;
  }
}

// v[k] = Vt[k][item]
__global__ void dpp_gather_v_kernel(const double* __restrict__ Vt, int64_t m, int64_t n, const int64_t* __restrict__ items, int64_t step, double* __restrict__ v) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < m) v[k] = Vt[k * n + items[step]];
}
// c[j] = Ft[j][:] . x for j < rows (one warp per row)
__global__ void dpp_rowdots_kernel(const double* __restrict__ Ft, int64_t rows, int64_t m, const double* __restrict__ x, double* __restrict__ c) {
  const int lane = threadIdx.x & 31;
  const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= rows) return;
  double acc = 0.0;
  for (int64_t k = lane; k < m; k += 32) acc = fma(Ft[j * m + k], x[k], acc);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) c[j] = acc;
}
// x[k] -= sum_{j < rows} c[j] Ft[j][k], in two deterministic stages: blockIdx.y slices the row range (a single stage had only m / 128
// blocks in flight and ran at a tenth of the HBM rate), partial sums go to tmp[slice][k]
__global__ void dpp_subtract_partial_kernel(const double* __restrict__ Ft, int64_t rows, int64_t m, const double* __restrict__ c, int64_t jslice,
                                            double* __restrict__ tmp) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const int64_t j0 = (int64_t)blockIdx.y * jslice;
  int64_t j1 = j0 + jslice;
  if (j1 > rows) j1 = rows;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int64_t j = j0;
  for (; j + 4 <= j1; j += 4) {
    s0 = fma(c[j], __ldg(Ft + j * m + k), s0);
    s1 = fma(c[j + 1], __ldg(Ft + (j + 1) * m + k), s1);
    s2 = fma(c[j + 2], __ldg(Ft + (j + 2) * m + k), s2);
    s3 = fma(c[j + 3], __ldg(Ft + (j + 3) * m + k), s3);
  }
  for (; j < j1; ++j) s0 = fma(c[j], __ldg(Ft + j * m + k), s0);
  tmp[(int64_t)blockIdx.y * m + k] = (s0 + s1) + (s2 + s3);
}
__global__ void dpp_subtract_finish_kernel(const double* __restrict__ tmp, int nslices, int64_t m, double* __restrict__ x) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  double t = 0.0;
  for (int s = 0; s < nslices; ++s) t += tmp[(int64_t)s * m + k];
  x[k] -= t;
}
// Ft[step][:] = x / |x|  (one block)
__global__ void dpp_normalise_kernel(const double* __restrict__ x, int64_t m, double* __restrict__ Frow) {
  __shared__ double sh[1024];
  double a = 0.0;
  for (int64_t k = threadIdx.x; k < m; k += blockDim.x) a = fma(x[k], x[k], a);
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  const double inv = sh[0] > 0.0 ? 1.0 / sqrt(sh[0]) : 0.0;
  for (int64_t k = threadIdx.x; k < m; k += blockDim.x) Frow[k] = x[k] * inv;
}
// p_j -= (sum_k Vt[k][j] f[k])^2, the picked item gets exactly 0. blockIdx.y slices the k range; slices > 0 accumulate into `tmp`
// with a second tiny kernel only when the range is split (nslices > 1).
__global__ void dpp_update_partial_kernel(const double* __restrict__ Vt, int64_t m, int64_t n, const double* __restrict__ f, int64_t kslice,
                                          double* __restrict__ tmp) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int64_t k0 = (int64_t)blockIdx.y * kslice;
  int64_t k1 = k0 + kslice;
  if (k1 > m) k1 = m;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int64_t k = k0;
  for (; k + 4 <= k1; k += 4) {
    s0 = fma(__ldg(Vt + k * n + j), f[k], s0);
    s1 = fma(__ldg(Vt + (k + 1) * n + j), f[k + 1], s1);
    s2 = fma(__ldg(Vt + (k + 2) * n + j), f[k + 2], s2);
    s3 = fma(__ldg(Vt + (k + 3) * n + j), f[k + 3], s3);
  }
  for (; k < k1; ++k) s0 = fma(__ldg(Vt + k * n + j), f[k], s0);
  tmp[(int64_t)blockIdx.y * n + j] = (s0 + s1) + (s2 + s3);
}
__global__ void dpp_update_finish_kernel(const double* __restrict__ tmp, int nslices, int64_t n, const int64_t* __restrict__ items, int64_t step,
                                         double* __restrict__ p) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double t = 0.0;
  for (int s = 0; s < nslices; ++s) t += tmp[(int64_t)s * n + j];
  double q = fma(-t, t, p[j]);
  if (q < 0.0 || j == items[step]) q = 0.0;
  p[j] = q;
}

// ---- the same draw in its kernel-row (Cholesky) form: used when the n x n projection kernel P = V V' fits on the device ----
// Conditioning on the picked items is the pivoted Cholesky recursion on P (Poulson 2019; Launay, Galerne & Desolneux 2018) — the
// same conditional distribution as the Gram-Schmidt form, item by item:
//   p = diag(P);  repeat m times:  item ~ p / sum(p);  c = (P[item,:] - C[:t,item]' C[:t,:]) / sqrt(p_item);  C[t,:] = c;  p -= c.^2.
// P comes out of the Gram engine (DMMA) once; a step then reads t rows of C instead of all of Vt plus four passes over F, and is three
// launches after the pick. V is the n x m matrix of the selected eigenvectors (row-major, leading dimension ldv, pad column zeroed).
__global__ void gather_cols_kernel(const double* __restrict__ U, int64_t n, int64_t ldu, const int64_t* __restrict__ idx, int64_t m, double* __restrict__ V,
                                   int64_t ldv) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  const int64_t j0 = (int64_t)blockIdx.x * 32, k0 = (int64_t)blockIdx.y * 32;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t k = k0 + ty + 8 * r, j = j0 + tx;
    tile[ty + 8 * r][tx] = (k < m && j < n) ? U[idx[k] * ldu + j] : 0.0;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t j = j0 + ty + 8 * r, k = k0 + tx;
    if (j < n && k < ldv) V[j * ldv + k] = tile[tx][ty + 8 * r];
  }
}
// cj[s] = C[s][item] for s < t, and the pivot p_item (read here so that the row update may overwrite p in place)
__global__ void dpp_chol_gather_kernel(const double* __restrict__ C, int64_t t, int64_t n, const int64_t* __restrict__ items, int64_t step,
                                       const double* __restrict__ p, double* __restrict__ cj, double* __restrict__ pivot) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t item = items[step];
  if (item < 0) return;
  if (s < t) cj[s] = C[s * n + item];
  if (s == 0) pivot[0] = p[item];
}
// c = (P[item][:] - sum of the partial slices) / sqrt(pivot);  C[t][:] = c;  p -= c^2 (clamped at 0; the picked item gets exactly 0)
__global__ void dpp_chol_finish_kernel(const double* __restrict__ P, int64_t n, int64_t ld, const double* __restrict__ tmp, int nslices,
                                       const int64_t* __restrict__ items, int64_t step, const double* __restrict__ pivot, double* __restrict__ Crow,
                                       double* __restrict__ p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t item = items[step];
  if (item < 0) return;
  double acc = 0.0;
  for (int s = 0; s < nslices; ++s) acc += tmp[(int64_t)s * n + i];
  const double c = (P[item * ld + i] - acc) / sqrt(pivot[0]);
  Crow[i] = c;
  double q = fma(-c, c, p[i]);
  if (q < 0.0 || i == item) q = 0.0;
  p[i] = q;
}

// ---- greedy MAP of the L-ensemble (Determinantal.greedy_subset, dpp.jl:60 [external]; Chen et al. 2018 fast greedy form) ----
//   d = diag(K);  repeat k times:  j = argmax d;  e = (K[j,:] - C[:t,j]' C[:t,:]) / sqrt(d_j);  C[t,:] = e;  d -= e.^2
// K: the kernel matrix kept on the device next to its eigenvectors (row-major lower triangle = Symmetric(:U) storage, leading dimension ld). One block picks the
// argmax (first index on ties, like argmax); the row update is the same two-stage deterministic reduction as the sampler's.
__global__ void greedy_init_kernel(const double* __restrict__ K, int64_t n, int64_t ld, double* __restrict__ dg) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) dg[j] = K[j * ld + j];
}
__global__ void greedy_argmax_kernel(const double* __restrict__ dg, int64_t n, int64_t step, int64_t* __restrict__ items) {
  __shared__ double bv[1024];
  __shared__ int64_t bi[1024];
  double best = -1.0e308;
  int64_t arg = -1;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
    const double x = dg[j];
    if (x > best) {  // strided scan keeps the smallest index among equal values within a thread
      best = x;
      arg = j;
    }
  }
  bv[threadIdx.x] = best;
  bi[threadIdx.x] = arg;
  __syncthreads();
  for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) {
      const double o = bv[threadIdx.x + off];
      const int64_t oi = bi[threadIdx.x + off];
      if (o > bv[threadIdx.x] || (o == bv[threadIdx.x] && oi >= 0 && (bi[threadIdx.x] < 0 || oi < bi[threadIdx.x]))) {
        bv[threadIdx.x] = o;
        bi[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) items[step] = bv[0] > 0.0 ? bi[0] : -1;  // a non-positive conditional variance: the ensemble's rank is exhausted
}
// cj[s] = C[s][item] for s < t
__global__ void greedy_gather_kernel(const double* __restrict__ C, int64_t t, int64_t n, const int64_t* __restrict__ items, int64_t step, double* __restrict__ cj) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < t && items[step] >= 0) cj[s] = C[s * n + items[step]];
}
// tmp[slice][i] = sum_{s in slice} cj[s] C[s][i]
__global__ void greedy_dot_partial_kernel(const double* __restrict__ C, int64_t t, int64_t n, const double* __restrict__ cj, int64_t sslice, double* __restrict__ tmp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t s0 = (int64_t)blockIdx.y * sslice;
  int64_t s1 = s0 + sslice;
  if (s1 > t) s1 = t;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int64_t s = s0;
  for (; s + 4 <= s1; s += 4) {
    a0 = fma(cj[s], __ldg(C + s * n + i), a0);
    a1 = fma(cj[s + 1], __ldg(C + (s + 1) * n + i), a1);
    a2 = fma(cj[s + 2], __ldg(C + (s + 2) * n + i), a2);
    a3 = fma(cj[s + 3], __ldg(C + (s + 3) * n + i), a3);
  }
  for (; s < s1; ++s) a0 = fma(cj[s], __ldg(C + s * n + i), a0);
  tmp[(int64_t)blockIdx.y * n + i] = (a0 + a1) + (a2 + a3);
}
// e = (K[item][:] - sum of the partial slices) / sqrt(d_item);  C[t][:] = e;  d -= e^2;  the picked item is retired (d = -inf)
__global__ void greedy_finish_kernel(const double* __restrict__ K, int64_t n, int64_t ld, const double* __restrict__ tmp, int nslices,
                                     const int64_t* __restrict__ items, int64_t step, double* __restrict__ Crow, double* __restrict__ dg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t item = items[step];
  if (item < 0) return;
  double acc = 0.0;
  for (int s = 0; s < nslices; ++s) acc += tmp[(int64_t)s * n + i];
  // K is held in the reference's Symmetric(:U) storage (row-major lower triangle): K[item][i] lives at [max][min]
  const double kji = i <= item ? K[item * ld + i] : K[i * ld + item];
  const double e = (kji - acc) / sqrt(dg[item]);
  Crow[i] = e;
  // (every thread reads dg[item] before any thread may overwrite it: the retire write is done by a separate launch)
  if (i != item) dg[i] = fma(-e, e, dg[i]);
}
__global__ void greedy_retire_kernel(const int64_t* __restrict__ items, int64_t step, double* __restrict__ dg) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && items[step] >= 0) dg[items[step]] = -1.0e308;
}

// ---- DBSCAN's O(n^2) distance matrices between configurations (src/SubsetSelection/dbscan.jl:100-170, kabsch.jl; row f4) ----
// P: n configurations x natoms x 3 positions, row-major. One thread per pair i < j; the positions (a few MB) live in L2.
// periodic_rmsd (dbscan.jl:110-124): d = p1 - p2; d -= round(d / L) * L (minimum image, round half to even); sum += norm(d)^2;
// sqrt(sum / natoms). Every operation is formed as the reference forms it (no fma contraction, norm then square), so the matrix is
// bit-identical to the restatement.
__global__ void rmsd_periodic_kernel(const double* __restrict__ P, int64_t n, int64_t natoms, double Lx, double Ly, double Lz, double* __restrict__ D,
                                     int64_t ldd) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (j >= n || j < i) return;
  if (j == i) {
    D[i * ldd + i] = 0.0;  // zeros(n, n): the loop runs over j > i only
    return;
  }
  const double* p1 = P + i * natoms * 3;
  const double* p2 = P + j * natoms * 3;
  const double L[3] = {Lx, Ly, Lz};
  double sum = 0.0;
  for (int64_t a = 0; a < natoms; ++a) {
    double ss = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double dd = __dsub_rn(p1[3 * a + c], p2[3 * a + c]);
      dd = __dsub_rn(dd, __dmul_rn(rint(__ddiv_rn(dd, L[c])), L[c]));
      ss = c == 0 ? __dmul_rn(dd, dd) : __dadd_rn(ss, __dmul_rn(dd, dd));
    }
    const double nr = sqrt(ss);  // norm(d)
    sum = __dadd_rn(sum, __dmul_rn(nr, nr));
  }
  const double r = sqrt(__ddiv_rn(sum, (double)natoms));
  D[i * ldd + j] = r;
  D[j * ldd + i] = r;
}

// centred positions Pc = P - centroid (translate_points, kabsch.jl:60-77), one thread per configuration
__global__ void centre_positions_kernel(const double* __restrict__ P, int64_t n, int64_t natoms, double* __restrict__ Pc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* p = P + i * natoms * 3;
  double c[3] = {0.0, 0.0, 0.0};
  for (int64_t a = 0; a < natoms; ++a)
    for (int k = 0; k < 3; ++k) c[k] += p[3 * a + k];
  for (int k = 0; k < 3; ++k) c[k] /= (double)natoms;
  for (int64_t a = 0; a < natoms; ++a)
    for (int k = 0; k < 3; ++k) Pc[i * natoms * 3 + 3 * a + k] = p[3 * a + k] - c[k];
}

// 3 x 3 SVD by one-sided Jacobi (Hestenes): rotates the columns of A until they are orthogonal, A V = U S.
__device__ __forceinline__ void svd3(const double (&A)[3][3], double (&U)[3][3], double (&S)[3], double (&V)[3][3]) {
  double B[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      B[r][c] = A[r][c];
      V[r][c] = r == c ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
        for (int r = 0; r < 3; ++r) {
          alpha += B[r][p] * B[r][p];
          beta += B[r][q] * B[r][q];
          gamma += B[r][p] * B[r][q];
        }
        if (gamma == 0.0) continue;
        off = fmax(off, fabs(gamma) / sqrt(alpha * beta + 1e-300));
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
        for (int r = 0; r < 3; ++r) {
          const double bp = B[r][p], bq = B[r][q];
          B[r][p] = cs * bp - sn * bq;
          B[r][q] = sn * bp + cs * bq;
          const double vp = V[r][p], vq = V[r][q];
          V[r][p] = cs * vp - sn * vq;
          V[r][q] = sn * vp + cs * vq;
        }
      }
    if (off < 1e-16) break;
  }
  for (int c = 0; c < 3; ++c) S[c] = sqrt(B[0][c] * B[0][c] + B[1][c] * B[1][c] + B[2][c] * B[2][c]);
  // order the singular values descending (only the completion of a vanishing direction below depends on it)
  for (int a = 0; a < 2; ++a)
    for (int b = a + 1; b < 3; ++b)
      if (S[b] > S[a]) {
        double tmp = S[a]; S[a] = S[b]; S[b] = tmp;
        for (int r = 0; r < 3; ++r) {
          tmp = B[r][a]; B[r][a] = B[r][b]; B[r][b] = tmp;
          tmp = V[r][a]; V[r][a] = V[r][b]; V[r][b] = tmp;
        }
      }
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) U[r][c] = S[c] > 0.0 ? B[r][c] / S[c] : 0.0;
  if (!(S[2] > 1e-14 * S[0])) {  // rank-deficient covariance (planar / collinear atoms): complete U with the cross product
    U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
    U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
    U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
  }
}
__device__ __forceinline__ double det3(const double (&M)[3][3]) {
  return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
         M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}
// kabsch_rmsd (kabsch.jl:91-124): A = Qc' Pc, (u, s, v) = svd(A), f = sign(det(v) det(u)), R = u diag(1,1,f) v', rmsd of Pc against Qc R.
__global__ void rmsd_kabsch_kernel(const double* __restrict__ Pc, int64_t n, int64_t natoms, double* __restrict__ D, int64_t ldd) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (j >= n || j < i) return;
  if (j == i) {
    D[i * ldd + i] = 0.0;
    return;
  }
  const double* p = Pc + i * natoms * 3;  // reference (P)
  const double* q = Pc + j * natoms * 3;  // coords (Q)
  double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int64_t a = 0; a < natoms; ++a)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) A[r][c] = fma(q[3 * a + r], p[3 * a + c], A[r][c]);
  double U[3][3], S[3], V[3][3];
  svd3(A, U, S, V);
  const double dd = det3(V) * det3(U);
  const double f = dd > 0.0 ? 1.0 : (dd < 0.0 ? -1.0 : 0.0);
  double R[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[r][c] = U[r][0] * V[c][0] + U[r][1] * V[c][1] + f * U[r][2] * V[c][2];
  double sum = 0.0;
  for (int64_t a = 0; a < natoms; ++a)
    for (int c = 0; c < 3; ++c) {
      const double rot = q[3 * a + 0] * R[0][c] + q[3 * a + 1] * R[1][c] + q[3 * a + 2] * R[2][c];
      const double e = p[3 * a + c] - rot;
      sum = fma(e, e, sum);
    }
  const double r = sqrt(sum / (double)natoms);
  D[i * ldd + j] = r;
  D[j * ldd + i] = r;
}

// ---- residual metrics of the fitted model (src/Metrics/metrics.jl:11-53) ----
// partial[b][0..6] = sums over block b's strided range of: |d|, d^2, x, (x - xmean)^2 (second pass), cos of 3-vector pairs, count of
// non-NaN cosines, with d = x_pred - x. Two passes (mean first: rsq needs sum (x - mean(x))^2 without cancellation). Block sums are
// added in block order by metrics_final_kernel -> deterministic.
__global__ void metrics_partial_kernel(const double* __restrict__ xp, const double* __restrict__ x, int64_t n, const double* __restrict__ xmean,
                                       double* __restrict__ partial) {
  __shared__ double sh[6][256];
  double a[6] = {0, 0, 0, 0, 0, 0};
  const double mu = xmean ? xmean[0] : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double d = xp[i] - x[i];
    a[0] += fabs(d);
    a[1] = fma(d, d, a[1]);
    a[2] += x[i];
    const double c = x[i] - mu;
    a[3] = fma(c, c, a[3]);
  }
  if (n % 3 == 0) {  // mean_cos over the 3-vectors (forces): dot(x, x_pred) / (|x| |x_pred|), NaNs (zero vectors) filtered
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n / 3; v += (int64_t)gridDim.x * blockDim.x) {
      const double a0 = x[3 * v], a1 = x[3 * v + 1], a2 = x[3 * v + 2];
      const double b0 = xp[3 * v], b1 = xp[3 * v + 1], b2 = xp[3 * v + 2];
      const double cs = (a0 * b0 + a1 * b1 + a2 * b2) / (sqrt(a0 * a0 + a1 * a1 + a2 * a2) * sqrt(b0 * b0 + b1 * b1 + b2 * b2));
      if (cs == cs) {
        a[4] += cs;
        a[5] += 1.0;
      }
    }
  }
  for (int q = 0; q < 6; ++q) sh[q][threadIdx.x] = a[q];
  __syncthreads();
  for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off)
      for (int q = 0; q < 6; ++q) sh[q][threadIdx.x] += sh[q][threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x < 6) partial[(int64_t)blockIdx.x * 6 + threadIdx.x] = sh[threadIdx.x][0];
}
// mode 0: out[0] = mean(x) (after the first pass). mode 1: out = [mae, rmse, rsq, mean_cos]
__global__ void metrics_final_kernel(const double* __restrict__ partial, int nblocks, int64_t n, int mode, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int b = 0; b < nblocks; ++b)
    for (int q = 0; q < 6; ++q) s[q] += partial[(int64_t)b * 6 + q];
  if (mode == 0) {
    out[0] = s[2] / (double)n;
  } else {
    out[0] = s[0] / (double)n;
    out[1] = sqrt(s[1] / (double)n);
    out[2] = 1.0 - s[1] / s[3];
    out[3] = s[4] / s[5];  // NaN when n is not a multiple of 3 or every cosine was NaN (mean of an empty collection)
  }
}

}  // namespace plb
