// eigh.cuh — hand-written symmetric eigensolver of libplb200 (row f1 of SURVEY.md 8f): eigen(Symmetric(Q)) of
// dimension_reduction.jl:13 (K x K covariance, K = 500 / 1000 in configs 3 / 4) and the eigendecomposition inside EllEnsemble(K)
// (dpp.jl:26,41; config 1: N = 1000). The reference calls LAPACK (syevr) here; round 1 called cuSOLVER. This file is the
// library-free path for matrices that live in L2 (n <= 4096; the default routing uses it up to n = 2048):
//
//   1. eigh_sytrd_*_kernel    Householder tridiagonalisation as ONE persistent cooperative kernel: the rank-2 update of step j and
//                             the symmetric matrix-vector product of step j+1 are one pass over the rows a warp owns (rows dealt
//                             cyclically; in registers for n <= 1184, through L2 above), every CTA forms the next reflector
//                             redundantly from the next row of the matrix (its next column: the row owners publish their
//                             elements), one grid barrier per step.
//   2. dc_*                   divide and conquer on the tridiagonal matrix (Cuppen / Gu-Eisenstat): Jacobi leaves (<= 32), per
//                             merge: deflation, secular equation by a bracketed rational iteration whose bracket lives ON THE BIT
//                             PATTERN of the offset from the nearer pole (a bisection step halves the number of doubles left,
//                             whatever the scale), Loewner re-computation of the update vector, and the eigenvector update
//                             Q <- Q V as a product on the DMMA Gram engine (gram.cuh, LINEAR epilogue).
//   3. eigh_backtransform_*   applies the reflectors to the eigenvectors of T, eight at a time in the compact WY form, as
//                             DMMA.8x8x4 with the eigenvectors in registers (n <= 2048; a scalar form and a one-warp-per-vector
//                             form remain for experiments and for larger n).
//
// scripts/eig_proto/dc_proto.py is the numpy model of step 2 (same formulas, same order) used to validate the algorithm on the CPU.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace plb {

constexpr int EIGH_LEAF = 32;          // largest leaf of the divide-and-conquer tree
constexpr int EIGH_MAX_N = 4096;       // shared-memory budgets below are sized for this
constexpr int EIGH_SYTRD_THREADS = 256;
constexpr double EIGH_EPS = 1.1102230246251565e-16;  // LAPACK dlamch('E')

struct EighMerge {
  int lo, mid, hi;  // node [lo, hi) with children [lo, mid) and [mid, hi)
};

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
__device__ __forceinline__ double warp_max(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}

// block-wide sum, result broadcast to every thread; `red` holds >= 33 doubles; two __syncthreads
__device__ __forceinline__ double block_sum(double x, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  x = warp_sum(x);
  __syncthreads();  // red may still be read from an earlier call
  if (lane == 0) red[wid] = x;
  __syncthreads();
  double t = lane < nw ? red[lane] : 0.0;
  return warp_sum(t);
}
__device__ __forceinline__ double block_max(double x, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  x = warp_max(x);
  __syncthreads();
  if (lane == 0) red[wid] = x;
  __syncthreads();
  double t = lane < nw ? red[lane] : 0.0;
  return warp_max(t);
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// ------------------------------------------------------------------------------------------------------------------------------
// max |a_ij| over the triangle that is read -> bits of the maximum in *out (atomicMax on the bit pattern of a non-negative double)
__global__ void eigh_maxabs_kernel(const double* __restrict__ A, int n, int64_t lda, int lower, unsigned long long* out) {
  double m = 0.0;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < (int64_t)n * n; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / n), j = (int)(e % n);
    if (lower ? j <= i : j >= i) {
      const double a = fabs(A[i * lda + j]);
      m = (a > m || a != a) ? a : m;  // a NaN wins
    }
  }
  // reduce the BIT PATTERNS (non-negative doubles order like their bits, NaN above Inf): fmax would drop a NaN
  unsigned long long mb = (unsigned long long)__double_as_longlong(fabs(m));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, mb, o);
    mb = t > mb ? t : mb;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(out, mb);
}

// S = full symmetric copy of the triangle of A, scaled by 1 / max|a_ij|; rowbuf0 = row 0 of S (the first "next row" of the sytrd loop)
__global__ void eigh_symmetrize_kernel(const double* __restrict__ A, int n, int64_t lda, int lower, const unsigned long long* maxbits,
                                       double* __restrict__ S, int64_t lds, double* __restrict__ rowbuf0, double* scal) {
  const double mx = __longlong_as_double((long long)*maxbits);
  const double inv = (mx > 0.0 && mx < INFINITY) ? 1.0 / mx : 1.0;
  if (blockIdx.x == 0 && threadIdx.x == 0) scal[0] = (mx > 0.0 && mx < INFINITY) ? mx : 1.0;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < (int64_t)n * n; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / n), j = (int)(e % n);
    const int hi = i > j ? i : j, lo = i > j ? j : i;
    const double a = (lower ? A[hi * lda + lo] : A[lo * lda + hi]) * inv;
    S[i * lds + j] = a;
    if (i == 0) rowbuf0[j] = a;
  }
}

// ------------------------------------------------------------------------------------------------------------------------------
struct SytrdParams {
  double* S;        // n x n full symmetric (row-major, lds); on return row j, columns > j: Householder vector j (v[j+1] = 1)
  int64_t lds;
  int n;
  double* d;        // n     diagonal of T
  double* e;        // n     off-diagonal of T (e[n-1] unused)
  double* tau;      // n
  double* p;        // 2 x n A v of the current step (rows written by their owners), double-buffered by step parity like rowbuf:
                    //       a CTA that is already forming the next product must not overwrite what a slower CTA still reads
  double* pd;       // 2 x 256 partial v . (A v) per CTA, same parity scheme
  double* rowbuf;   // 2 x n the next row of the updated matrix (= its next column: every row owner publishes its element),
                    //       double-buffered by step parity
  unsigned long long* bar;  // [0, grid) one slot per CTA (zeroed by the host), [256] abort flag
  long long* prof;  // NULL, or 4 cycle counters per CTA (thread 0: staging, compute, publish, barrier) -- scripts/eigh_probe.py
};

// all CTAs are co-resident (cooperative launch). Grid barrier: fence + atomic add on one counter, thread 0 polls it. Measured against
// the alternatives on 148 CTAs x 256 threads (scripts/native/gridbar_probe.cu, profiles/r02e_gridbar_probe.json; cycles per publish +
// barrier + consume): this 3565, red.release counter 4001, eight sub-counters 4298, per-CTA flags 128 B apart polled by one warp of
// every CTA 4946, slots collected by one CTA + a go word 5700, packed flags 12298. bar[0] counter (zeroed by the host), bar[1028]
// abort flag: a barrier that is not passed within ~4 s sets it and every CTA leaves.
constexpr int EIGH_BAR_GO = 1024, EIGH_BAR_ABORT = 1028, EIGH_BAR_WORDS = 1032;
__device__ __forceinline__ bool sytrd_grid_barrier(unsigned long long* bar, unsigned long long step) {
  __shared__ int s_ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long want = step * gridDim.x;
    __threadfence();
    atomicAdd(bar, 1ull);
    int ok = 1;
    const long long t0 = clock64();
    for (unsigned it = 1; ld_acquire_u64(bar) < want; ++it) {
      if ((it & 1023u) == 0) {
        if (ld_acquire_u64(bar + EIGH_BAR_ABORT) != 0ull) { ok = 0; break; }
        if (clock64() - t0 > 8000000000ll) { atomicExch(bar + EIGH_BAR_ABORT, 1ull); ok = 0; break; }
      }
    }
    s_ok = ok;
  }
  __syncthreads();
  return s_ok != 0;
}

// n <= 8 * gridDim.x: every warp owns ONE row and keeps it in REGISTERS for the whole reduction (lane l holds columns l, l+32, ...); the
// matrix never travels again. Per step: (1) the CTA forms w = tau p - alpha v and the next row of the updated matrix element-wise in
// shared memory from the products / broadcast row the owners published before the barrier (every element once per CTA); (2) every
// warp updates its row, forms its entry of the next product and sums the norm of the next reflector in the same pass -- a BRANCH-FREE
// unrolled loop (finished columns are updated with zeros / ignored rather than skipped: a 25-cycle DFMA latency wants independent
// chunks in flight, and a branch per chunk serialises them: 2500 -> 600 cycles); (3) publish, grid barrier.
// Nothing is spilled: acquire loads invalidate L1 every step, so a spilled register costs an L2 round trip per reload.
template <int NCH, int START, bool LIVE, bool SIG = true>
__device__ __forceinline__ void sytrd_row_pass(double (&a)[NCH], const double* __restrict__ svl, const double* __restrict__ swl,
                                               const double* __restrict__ srnl, double nvr, double nwr, int mq, double (&acc)[4], double (&sg)[4],
                                               double& cap) {
#pragma unroll
  for (int m = START; m < NCH; ++m) {
    const double rn = srnl[32 * m];
    if (SIG) sg[m & 3] = fma(rn, rn, sg[m & 3]);
    if (LIVE) {
      const double an = fma(nwr, svl[32 * m], fma(nvr, swl[32 * m], a[m]));
      a[m] = an;
      acc[m & 3] = fma(an, rn, acc[m & 3]);
      cap = m == mq ? an : cap;
    }
  }
}
template <int NCH, bool LIVE>
__device__ __forceinline__ void sytrd_row_pass_from(int m0, double (&a)[NCH], const double* svl, const double* swl, const double* srnl, double nvr,
                                                    double nwr, int mq, double (&acc)[4], double (&sg)[4], double& cap) {
  switch (m0 >> 3) {  // finished chunks are skipped eight at a time only: no branch inside the unrolled pass
    case 0: sytrd_row_pass<NCH, 0, LIVE>(a, svl, swl, srnl, nvr, nwr, mq, acc, sg, cap); break;
    case 1: sytrd_row_pass<NCH, (8 < NCH ? 8 : 0), LIVE>(a, svl, swl, srnl, nvr, nwr, mq, acc, sg, cap); break;
    case 2: sytrd_row_pass<NCH, (16 < NCH ? 16 : 0), LIVE>(a, svl, swl, srnl, nvr, nwr, mq, acc, sg, cap); break;
    case 3: sytrd_row_pass<NCH, (24 < NCH ? 24 : 0), LIVE>(a, svl, swl, srnl, nvr, nwr, mq, acc, sg, cap); break;
    default: sytrd_row_pass<NCH, (32 < NCH ? 32 : 0), LIVE>(a, svl, swl, srnl, nvr, nwr, mq, acc, sg, cap); break;
  }
}

template <int NCH>
__global__ void __launch_bounds__(EIGH_SYTRD_THREADS, 1) eigh_sytrd_reg_kernel(SytrdParams P) {
  constexpr int NP = NCH * 32;
  constexpr int NQ = (NP + EIGH_SYTRD_THREADS - 1) / EIGH_SYTRD_THREADS;
  // srn: row jn of the updated matrix from column jn + 2 on (zero before: columns <= jn + 1 of every earlier step were written as zero
  // and are never touched again, so no pass needs a lower bound inside its first chunk); its entries jn and jn + 1 travel in s_sc
  __shared__ double svb[2][NP], sw[NP], srn[NP], s_part[8], s_sc[4];
  const int n = P.n;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int G = gridDim.x, b = blockIdx.x;
  const int r = b + G * wid;  // this warp's row (>= n: no row)
  double a[NCH];
#pragma unroll
  for (int m = 0; m < NCH; ++m) {
    const int c = lane + 32 * m;
    a[m] = (r < n && c < n) ? __ldcg(P.S + (size_t)r * P.lds + c) : 0.0;
  }
  for (int c = tid; c < NP; c += EIGH_SYTRD_THREADS) { svb[0][c] = 0.0; svb[1][c] = 0.0; sw[c] = 0.0; srn[c] = 0.0; }
  __syncthreads();
  double tau = 0.0;
  int cur = 0;  // svb[cur] = v_j (zero outside its support), svb[cur ^ 1] <- v_{j+1}
  unsigned long long target = 0;
  long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tk = clock64();  // phase cycle counters of one warp (PLB200_EIGH_PROF)
#define PLB_PF(i) do { if (P.prof) { const long long t_ = clock64(); pf[i] += t_ - tk; tk = t_; } } while (0)
  for (int j = -1; j <= n - 2; ++j, cur ^= 1) {
    const int jn = j + 1;
    const int m0 = jn >> 5;
    const int cbase = (jn & ~31) + tid;
    const double* sv = svb[cur];
    double* svn = svb[cur ^ 1];
    // ---- (1) w_j and row jn of the updated matrix, element-wise
    {
      const double* rb = P.rowbuf + (size_t)((j - 1) & 1) * n;
      const double* pin = P.p + (size_t)((j - 1) & 1) * n;
      const double* pdin = P.pd + ((j - 1) & 1) * 256;
      double pc[NQ], rbc[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int c = cbase + q * EIGH_SYTRD_THREADS;
        pc[q] = (c < n && j >= 0) ? __ldcg(pin + c) : 0.0;
        rbc[q] = c < n ? __ldcg(rb + c) : 0.0;
      }
      // the partial dots and p[jn] are fetched by ONE warp per CTA (every thread asking for the same few lines is 6000 requests per
      // step on ten L2 lines) and handed on through shared memory
      if (wid == 0) {
        double dot = 0.0, pjn = 0.0;
        if (j >= 0) {
#pragma unroll
          for (int q = 0; q < 5; ++q) dot += __ldcg(pdin + lane + 32 * q);
          pjn = __ldcg(pin + jn);
        }
        dot = warp_sum(dot);
        if (lane == 0) {
          const double al = 0.5 * tau * tau * dot;
          s_sc[0] = al;
          s_sc[1] = tau * pjn - al;  // w_j[jn]; v_j[jn] = 1
        }
      }
      __syncthreads();
      const double alpha = s_sc[0], wjn = s_sc[1];
      PLB_PF(0);  // loads of p, row, partial dots returned
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int c = cbase + q * EIGH_SYTRD_THREADS;
        if (c >= jn && c < n) {
          const double vc = sv[c];
          const double wc = tau * pc[q] - alpha * vc;
          const double rn = rbc[q] - (wc + wjn * vc);
          sw[c] = wc;
          srn[c] = c >= jn + 2 ? rn : 0.0;
          if (c == jn) s_sc[2] = rn;
          if (c == jn + 1) s_sc[3] = rn;
        }
      }
      PLB_PF(1);  // element-wise w and next row
    }
    __syncthreads();
    PLB_PF(2);  // CTA barrier
    const double djn = s_sc[2];
    if (jn == n - 1) {
      if (b == 0 && tid == 0) P.d[jn] = djn;
      break;
    }
    const double a0 = s_sc[3];
    // ---- (2) one pass over the own row: rank-2 update with (v_j, w_j), the row's entry of A v_{jn} (unscaled: the reflector's
    // scalars are not known yet -- they need sigma, which every warp sums in this same pass), and the entry in column jn + 1
    const bool live = r > jn && r < n;
    double acc[4] = {0.0, 0.0, 0.0, 0.0}, sg[4] = {0.0, 0.0, 0.0, 0.0}, cap = 0.0;
    const int mq = (jn + 1) >> 5;
    if (live) sytrd_row_pass_from<NCH, true>(m0, a, sv + lane, sw + lane, srn + lane, -sv[r], -sw[r], mq, acc, sg, cap);
    else sytrd_row_pass_from<NCH, false>(m0, a, sv + lane, sw + lane, srn + lane, 0.0, 0.0, mq, acc, sg, cap);
    double sigma = (sg[0] + sg[1]) + (sg[2] + sg[3]), accs = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {  // two reductions, interleaved
      sigma += __shfl_xor_sync(0xffffffffu, sigma, o);
      accs += __shfl_xor_sync(0xffffffffu, accs, o);
    }
    const double acc0 = __shfl_sync(0xffffffffu, cap, (jn + 1) & 31);
    double beta = a0, taun = 0.0, scl = 0.0;
    if (sigma > 0.0) {
      // beta = -sign(a0) sqrt(x), tau = (beta - a0) / beta = 1 + |a0| / sqrt(x), 1 / (a0 - beta) = sign(a0) / (|a0| + sqrt(x)):
      // rsqrt + one reciprocal instead of sqrt + two divisions on the critical path of every step (124-cycle divisions)
      const double x = fma(a0, a0, sigma), y = rsqrt(x), sx = x * y, aa = fabs(a0);
      beta = -copysign(sx, a0);
      taun = fma(aa, y, 1.0);
      scl = copysign(1.0 / (aa + sx), a0);
    }
    double pdw = 0.0;
    if (live) {
      const double prod = fma(scl, accs, acc0);
      const double vnr = r >= jn + 2 ? srn[r] * scl : 1.0;
      pdw = vnr * prod;
      if (lane == 0) {
        __stcg(P.p + (size_t)(j & 1) * n + r, prod);
        // row jn + 1 of the updated matrix, which every CTA needs next, is column jn + 1 by symmetry: every row owner publishes
        // its own element of it (no owner of the whole row on the critical path of the step)
        __stcg(P.rowbuf + (size_t)(j & 1) * n + r, acc0);
      }
    }
    if (lane == 0) s_part[wid] = pdw;
    PLB_PF(3);  // row pass, scalars, product entry, broadcast row
    __syncthreads();  // every warp is done with sv / sw / srn of this step
    // ---- (3) the scaled reflector into the other buffer (the owner of row jn also stores it), T's entries, partial dot
    {
      double* row = P.S + (size_t)jn * P.lds;
      const bool owner = (jn % G == b);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int c = cbase + q * EIGH_SYTRD_THREADS;
        if (c < NP) {
          const double x = (c >= jn + 2 && c < n) ? srn[c] * scl : (c == jn + 1 ? 1.0 : 0.0);
          svn[c] = x;
          if (owner && c > jn && c < n) row[c] = x;
        }
      }
      if (tid == 0) {
        if (jn >= 1) svn[jn - 1] = 0.0;  // the leading 1 of v_{j-1}; everything before it is zero already
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += s_part[q];
        __stcg(P.pd + (j & 1) * 256 + b, t);
        if (b == 0) { P.d[jn] = djn; P.e[jn] = beta; P.tau[jn] = taun; }
      }
    }
    tau = taun;
    target += 1ull;
    PLB_PF(4);  // reflector in place + publish
    if (!sytrd_grid_barrier(P.bar, target)) return;
    PLB_PF(5);  // grid barrier
  }
  if (P.prof && lane == 0 && wid == 6) {
    for (int i = 0; i < 8; ++i) P.prof[8 * b + i] = pf[i];
  }
#undef PLB_PF
}

// The same step structure for 1184 < n <= 1024 NB with the rows in L2: 12 warps per CTA, a warp's rows (dealt cyclically, one per
// warp up to n = 1776) pass through its registers one block of 32 chunks (1024 columns) at a time -- 32 independent loads in flight per
// lane, the fused pass, predicated stores -- and the other warps of the SM hide the round trip. Shared memory holds v_j, v_{j+1}, w_j
// and the next row (dynamic: 4 x 8 NB KB). SMR (n <= 2048): the columns >= 1024 of the CTA's rows -- the part of the matrix that lives
// longest -- stay in shared memory (<= 14 rows x 8 KB): a third of the L2 traffic (67 MB per step at n = 2048 without it). Measured and
// rejected: ALL of every row in shared memory (n <= 1536) is slower (n = 1500: 9.4 -> 10.5 ms) -- the pass then moves five 256-byte
// shared-memory transactions per 32 columns and row (v, w, next row, row in, row out) against 128 B / clk / SM; prefetching the next
// step's first block across the grid barrier changes nothing.
constexpr int EIGH_L2_THREADS = 384;  // 12 warps = 3 per scheduler: 168 registers per thread keep the 32-chunk block and the pass without spills
constexpr int EIGH_L2_WARPS = EIGH_L2_THREADS / 32;
template <bool LIVE, bool SIG>
__device__ __forceinline__ void sytrd_block_pass(int start8, double (&a)[32], const double* svl, const double* swl, const double* srnl, double nvr,
                                                 double nwr, int mq, double (&acc)[4], double (&sg)[4], double& cap) {
  switch (start8) {
    case 0: sytrd_row_pass<32, 0, LIVE, SIG>(a, svl, swl, srnl, nvr, nwr, mq, acc, sg, cap); break;
    case 1: sytrd_row_pass<32, 8, LIVE, SIG>(a, svl, swl, srnl, nvr, nwr, mq, acc, sg, cap); break;
    case 2: sytrd_row_pass<32, 16, LIVE, SIG>(a, svl, swl, srnl, nvr, nwr, mq, acc, sg, cap); break;
    default: sytrd_row_pass<32, 24, LIVE, SIG>(a, svl, swl, srnl, nvr, nwr, mq, acc, sg, cap); break;
  }
}

template <int NB, bool SMR>
__global__ void __launch_bounds__(EIGH_L2_THREADS, 1) eigh_sytrd_l2_kernel(SytrdParams P) {
  constexpr int NP = NB * 1024;
  constexpr int NQ = (NP + EIGH_L2_THREADS - 1) / EIGH_L2_THREADS;
  extern __shared__ double l2sm[];
  double* svb0 = l2sm;
  double* svb1 = l2sm + NP;
  double* sw = l2sm + 2 * NP;
  double* srn = l2sm + 3 * NP;
  double* rowsm = l2sm + 4 * NP;  // SMR: columns >= 1024 of the CTA's rows (slot (r - b) / G), 1024 doubles each
  __shared__ double s_part[EIGH_L2_WARPS], s_sc[4];
  const int n = P.n;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int G = gridDim.x, b = blockIdx.x;
  for (int c = tid; c < NP; c += EIGH_L2_THREADS) { svb0[c] = 0.0; svb1[c] = 0.0; sw[c] = 0.0; srn[c] = 0.0; }
  if (SMR) {
    for (int rr = b + G * wid, slot = wid; rr < n; rr += G * EIGH_L2_WARPS, slot += EIGH_L2_WARPS)
      for (int c = lane; c < 1024; c += 32) rowsm[slot * 1024 + c] = 1024 + c < n ? __ldcg(P.S + (size_t)rr * P.lds + 1024 + c) : 0.0;
  }
  __syncthreads();
  double tau = 0.0;
  int cur = 0;
  unsigned long long target = 0;
  long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tk = clock64();  // phase cycle counters of one warp (PLB200_EIGH_PROF)
#define PLB_PF(i) do { if (P.prof) { const long long t_ = clock64(); pf[i] += t_ - tk; tk = t_; } } while (0)
  for (int j = -1; j <= n - 2; ++j, cur ^= 1) {
    const int jn = j + 1;
    const int m0 = jn >> 5;
    const int cbase = (jn & ~31) + tid;
    const double* sv = cur ? svb1 : svb0;
    double* svn = cur ? svb0 : svb1;
    // ---- (1) w_j and row jn of the updated matrix, element-wise
    {
      const double* rb = P.rowbuf + (size_t)((j - 1) & 1) * n;
      const double* pin = P.p + (size_t)((j - 1) & 1) * n;
      const double* pdin = P.pd + ((j - 1) & 1) * 256;
      double pc[NQ], rbc[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int c = cbase + q * EIGH_L2_THREADS;
        pc[q] = (c < n && j >= 0) ? __ldcg(pin + c) : 0.0;
        rbc[q] = c < n ? __ldcg(rb + c) : 0.0;
      }
      if (wid == 0) {
        double dot = 0.0, pjn = 0.0;
        if (j >= 0) {
#pragma unroll
          for (int q = 0; q < 5; ++q) dot += __ldcg(pdin + lane + 32 * q);
          pjn = __ldcg(pin + jn);
        }
        dot = warp_sum(dot);
        if (lane == 0) {
          const double al = 0.5 * tau * tau * dot;
          s_sc[0] = al;
          s_sc[1] = tau * pjn - al;
        }
      }
      __syncthreads();
      const double alpha = s_sc[0], wjn = s_sc[1];
      PLB_PF(0);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int c = cbase + q * EIGH_L2_THREADS;
        if (c >= jn && c < n) {
          const double vc = sv[c];
          const double wc = tau * pc[q] - alpha * vc;
          const double rn = rbc[q] - (wc + wjn * vc);
          sw[c] = wc;
          srn[c] = c >= jn + 2 ? rn : 0.0;
          if (c == jn) s_sc[2] = rn;
          if (c == jn + 1) s_sc[3] = rn;
        }
      }
    }
    PLB_PF(1);
    __syncthreads();
    PLB_PF(2);
    const double djn = s_sc[2];
    if (jn == n - 1) {
      if (b == 0 && tid == 0) P.d[jn] = djn;
      break;
    }
    const double a0 = s_sc[3];
    // ---- (2) the warp's rows, block by block; the first pass of the step also sums the norm of the next reflector
    const int mq = (jn + 1) >> 5;
    const int kb0 = m0 >> 5;  // first block with live columns
    double sg[4] = {0.0, 0.0, 0.0, 0.0};
    double sigma = 0.0, beta = a0, taun = 0.0, scl = 0.0, pdw = 0.0;
    bool have_sigma = false;
    int r = b + G * wid;
    if (r <= jn) r += ((jn - r) / (EIGH_L2_WARPS * G) + 1) * (EIGH_L2_WARPS * G);  // first live row of this warp
    for (int pass = 0; r < n || !have_sigma; ++pass, r += EIGH_L2_WARPS * G) {
      const bool live = r < n;
      double acc[4] = {0.0, 0.0, 0.0, 0.0}, cap = 0.0;
      const double nvr = live ? -sv[r] : 0.0, nwr = live ? -sw[r] : 0.0;
      double* row = P.S + (size_t)(live ? r : 0) * P.lds;
      double* rbn = P.rowbuf + (size_t)(j & 1) * n;
      for (int kb = kb0; kb < NB; ++kb) {
        if (kb * 1024 >= n) break;
        const int off = kb * 1024 + lane;
        const int rel = m0 - kb * 32;
        const int start8 = rel <= 0 ? 0 : (rel >> 3);
        double a[32];
        if (live) {
          double* srow = rowsm + (size_t)((r - b) / G) * 1024 + lane;  // SMR, kb == 1: this row's block lives in shared memory
          if (SMR && kb == 1) {
#pragma unroll
            for (int m = 0; m < 32; ++m) a[m] = srow[32 * m];
          } else {
#pragma unroll
            for (int m = 0; m < 32; ++m) {
              const int c = off + 32 * m;
              a[m] = c < n ? __ldcg(row + c) : 0.0;
            }
          }
          if (!have_sigma) sytrd_block_pass<true, true>(start8, a, sv + off, sw + off, srn + off, nvr, nwr, mq - kb * 32, acc, sg, cap);
          else sytrd_block_pass<true, false>(start8, a, sv + off, sw + off, srn + off, nvr, nwr, mq - kb * 32, acc, sg, cap);
          if (SMR && kb == 1) {
#pragma unroll
            for (int m = 0; m < 32; ++m) {
              const int c = off + 32 * m;
              if (c > jn && c < n) {
                srow[32 * m] = a[m];
              }
            }
          } else {
#pragma unroll
            for (int m = 0; m < 32; ++m) {
              const int c = off + 32 * m;
              if (c > jn && c < n) {
                __stcg(row + c, a[m]);
              }
            }
          }
        } else {
          sytrd_block_pass<false, true>(start8, a, sv + off, sw + off, srn + off, 0.0, 0.0, -1, acc, sg, cap);
        }
      }
      double accs = (acc[0] + acc[1]) + (acc[2] + acc[3]);
      if (!have_sigma) {
        sigma = (sg[0] + sg[1]) + (sg[2] + sg[3]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          sigma += __shfl_xor_sync(0xffffffffu, sigma, o);
          accs += __shfl_xor_sync(0xffffffffu, accs, o);
        }
        if (sigma > 0.0) {
          const double x = fma(a0, a0, sigma), y = rsqrt(x), sx = x * y, aa = fabs(a0);
          beta = -copysign(sx, a0);
          taun = fma(aa, y, 1.0);
          scl = copysign(1.0 / (aa + sx), a0);
        }
        have_sigma = true;
      } else {
        accs = warp_sum(accs);
      }
      if (live) {
        const double acc0 = __shfl_sync(0xffffffffu, cap, (jn + 1) & 31);
        const double prod = fma(scl, accs, acc0);
        const double vnr = r >= jn + 2 ? srn[r] * scl : 1.0;
        pdw += vnr * prod;
        if (lane == 0) {
          __stcg(P.p + (size_t)(j & 1) * n + r, prod);
          __stcg(rbn + r, acc0);  // this row's element of the next row (column jn + 1 by symmetry)
        }
      }
    }
    if (lane == 0) s_part[wid] = pdw;
    PLB_PF(3);
    __syncthreads();
    // ---- (3) the scaled reflector into the other buffer (the owner of row jn also stores it), T's entries, partial dot
    {
      double* row = P.S + (size_t)jn * P.lds;
      const bool owner = (jn % G == b);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int c = cbase + q * EIGH_L2_THREADS;
        if (c < NP) {
          const double x = (c >= jn + 2 && c < n) ? srn[c] * scl : (c == jn + 1 ? 1.0 : 0.0);
          svn[c] = x;
          if (owner && c > jn && c < n) row[c] = x;
        }
      }
      if (tid == 0) {
        if (jn >= 1) svn[jn - 1] = 0.0;
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < EIGH_L2_WARPS; ++q) t += s_part[q];
        __stcg(P.pd + (j & 1) * 256 + b, t);
        if (b == 0) { P.d[jn] = djn; P.e[jn] = beta; P.tau[jn] = taun; }
      }
    }
    tau = taun;
    target += 1ull;
    PLB_PF(4);
    if (!sytrd_grid_barrier(P.bar, target)) return;
    PLB_PF(5);
  }
  if (P.prof && lane == 0 && wid == 6) {
    for (int i = 0; i < 8; ++i) P.prof[8 * b + i] = pf[i];
  }
#undef PLB_PF
}

// ------------------------------------------------------------------------------------------------------------------------------
// divide and conquer on T = tridiag(d, e)

// leaves: cyclic Jacobi (round-robin ordering, 16 disjoint rotations per round) on the <= 32 x 32 diagonal block, torn at both ends.
// One CTA of 256 threads per leaf: thread (x, y) owns the 2 x 2 block between rotation pair x and rotation pair y and applies both
// sides of the round in one pass from one copy of A into the other (two __syncthreads per round). Also (every CTA, redundantly) the
// scale max(|d|, |e|) of T; scal[1] is written by leaf 0.
__global__ void __launch_bounds__(256) dc_leaf_kernel(const double* __restrict__ d, const double* __restrict__ e, int n, const int2* __restrict__ leaves,
                                                      double* __restrict__ lam, double* __restrict__ Q, int64_t ldq, double* scal) {
  __shared__ double Ab[2][32][33];
  __shared__ double V[32][33];
  __shared__ double cs[16][2];
  __shared__ int prs[16][2];
  __shared__ double red[40];
  const int tid = threadIdx.x;
  const int lo = leaves[blockIdx.x].x, hi = leaves[blockIdx.x].y, m = hi - lo;
  double mx = 0.0;
  for (int i = tid; i < n; i += blockDim.x) {
    mx = fmax(mx, fabs(d[i]));
    if (i + 1 < n) mx = fmax(mx, fabs(e[i]));
  }
  mx = block_max(mx, red);
  const double scale = mx > 0.0 ? mx : 1.0;
  const double inv = 1.0 / scale;
  if (blockIdx.x == 0 && tid == 0) scal[1] = scale;
  for (int t = tid; t < 32 * 32; t += blockDim.x) {
    const int i = t >> 5, j = t & 31;
    double a = 0.0;
    if (i < m && j < m) {
      if (i == j) {
        a = d[lo + i] * inv;
        if (i == 0 && lo > 0) a -= fabs(e[lo - 1] * inv);
        if (i == m - 1 && hi < n) a -= fabs(e[hi - 1] * inv);
      } else if (j == i + 1) a = e[lo + i] * inv;
      else if (i == j + 1) a = e[lo + j] * inv;
    }
    Ab[0][i][j] = a;
    V[i][j] = i == j ? 1.0 : 0.0;
  }
  __syncthreads();
  int cur = 0;
  const int kx = tid >> 4, ky = tid & 15;
  for (int sweep = 0; sweep < 24; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int t = tid; t < 32 * 32; t += blockDim.x) {
      const int i = t >> 5, j = t & 31;
      const double a = Ab[cur][i][j];
      if (i == j) dg += a * a; else off += a * a;
    }
    off = block_sum(off, red);
    dg = block_sum(dg, red);
    // converged: ||off||_F <= 1e-15 ||diag||_F (rotations leave rounding noise of that size behind: a tighter test never fires and the
    // kernel runs all 24 sweeps, measured: 400 us instead of 150)
    if (off <= 1e-30 * dg || off == 0.0) break;
    const double skip = 1e-36 * (dg + off);  // a_pq^2 below this is noise
    for (int r = 0; r < 31; ++r) {
      if (tid < 16) {
        int p, q;
        if (tid == 0) { p = r; q = 31; }
        else { p = r + tid; p = p >= 31 ? p - 31 : p; q = r - tid; q = q < 0 ? q + 31 : q; }
        if (p > q) { const int t = p; p = q; q = t; }
        const double apq = Ab[cur][p][q];
        double c = 1.0, s = 0.0;
        if (apq * apq > skip) {
          const double zeta = Ab[cur][q][q] - Ab[cur][p][p];
          const double t = (zeta >= 0.0 ? 2.0 : -2.0) * apq / (fabs(zeta) + sqrt(zeta * zeta + 4.0 * apq * apq));
          c = rsqrt(t * t + 1.0);
          s = t * c;
        }
        cs[tid][0] = c; cs[tid][1] = s;
        prs[tid][0] = p; prs[tid][1] = q;
      }
      __syncthreads();
      {
        // A' = J' A J on the 2 x 2 block (rows of pair kx) x (columns of pair ky)
        const double cx = cs[kx][0], sx = cs[kx][1], cy = cs[ky][0], sy = cs[ky][1];
        const int px = prs[kx][0], qx = prs[kx][1], py = prs[ky][0], qy = prs[ky][1];
        const double app = Ab[cur][px][py], apq = Ab[cur][px][qy], aqp = Ab[cur][qx][py], aqq = Ab[cur][qx][qy];
        // columns: (.,py) <- cy (.,py) - sy (.,qy) ; (.,qy) <- sy (.,py) + cy (.,qy)
        const double bpp = cy * app - sy * apq, bpq = sy * app + cy * apq;
        const double bqp = cy * aqp - sy * aqq, bqq = sy * aqp + cy * aqq;
        // rows: (px,.) <- cx (px,.) - sx (qx,.) ; (qx,.) <- sx (px,.) + cx (qx,.)
        Ab[cur ^ 1][px][py] = cx * bpp - sx * bqp;
        Ab[cur ^ 1][px][qy] = cx * bpq - sx * bqq;
        Ab[cur ^ 1][qx][py] = sx * bpp + cx * bqp;
        Ab[cur ^ 1][qx][qy] = sx * bpq + cx * bqq;
        // V <- V J: 16 pairs x 32 rows, two items per thread
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int it = tid + 256 * h;
          const int k = it >> 5, i = it & 31;
          const double c = cs[k][0], s2 = cs[k][1];
          if (s2 != 0.0) {
            const int p = prs[k][0], q = prs[k][1];
            const double vp = V[i][p], vq = V[i][q];
            V[i][p] = c * vp - s2 * vq;
            V[i][q] = s2 * vp + c * vq;
          }
        }
      }
      cur ^= 1;
      __syncthreads();
    }
  }
  for (int t = tid; t < 32 * 32; t += blockDim.x) {
    const int i = t >> 5, j = t & 31;
    if (i < m && j < m) Q[(size_t)(lo + i) * ldq + lo + j] = V[i][j];
  }
  if (tid < m) lam[lo + tid] = Ab[cur][tid][tid];
}

// per merge: z, sort, deflation (dlaed2's criteria). One CTA per merge.
// Outputs (all indexed from lo): perm[c'] = old column of new column c' (non-deflated ascending, then deflated), rot = (a, b, c, s)
// column rotations in order, Kp = number of non-deflated, dl / wz = poles and weights of the secular equation, lam = eigenvalues in the
// new order (the first Kp are overwritten by the secular kernel).
struct DeflateParams {
  const EighMerge* merges;
  const double* e;
  const double* scal;     // scal[1] = scale of T
  double* lam;
  const double* Q;
  int64_t ldq;
  int* perm;
  double4* rot;
  int* nrot;
  int* Kp;
  double* rho;
  double* dl;
  double* wz;
};

// exclusive prefix sum of one int per thread over the block (blockDim a multiple of 32, <= 1024); *total = the sum. sw: 33 ints.
__device__ __forceinline__ int block_excl_scan(int v, int* sw, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  __syncthreads();  // sw may still be read from an earlier call
  if (lane == 31) sw[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int t = lane < nw ? sw[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    sw[lane] = t;  // inclusive over warps
  }
  __syncthreads();
  *total = sw[nw - 1];
  return x - v + (wid > 0 ? sw[wid - 1] : 0);
}

// Dynamic shared memory: 48 k bytes. Everything but the chained close-pole deflations runs in parallel: the weights that vanish
// (rho |z| <= tol) are found and compacted away by all threads, the close-pole test is evaluated for every neighbouring pair of
// survivors with their ORIGINAL values, and one thread then walks only the runs that start at a flagged pair (inside a run the
// rotated values feed the next test, exactly as dlaed2's sequential scan does; after a failed test the state is the original one again).
__global__ void __launch_bounds__(1024) dc_deflate_kernel(DeflateParams P) {
  extern __shared__ double dsm[];
  __shared__ double red[40];
  __shared__ int sw[33];
  const EighMerge mg = P.merges[blockIdx.x];
  const int lo = mg.lo, mid = mg.mid, k = mg.hi - mg.lo, tid = threadIdx.x, nt = blockDim.x;
  double* dd = dsm;            // original order
  double* z = dsm + k;
  double* sd = dsm + 2 * k;    // ascending order
  double* sz = dsm + 3 * k;
  int* order = (int*)(dsm + 4 * k);  // order[t] = original index of sorted position t
  int* surv = order + k;             // sorted positions of the weights that do not vanish
  int* flag = surv + k;              // per survivor: 1 = the pair (s-1, s) deflates with original values; later 2 = deflated by a rotation
  int* outp = flag + k;              // new column -> sorted position
  const double beta = P.e[mid - 1] * (1.0 / P.scal[1]);  // as the leaves scale
  const double rho = 2.0 * fabs(beta);
  const double sgn = beta < 0.0 ? -1.0 : 1.0;
  const double rs = 0.70710678118654752440;
  double mx = 0.0;
  for (int i = tid; i < k; i += nt) {
    const double di = P.lam[lo + i];
    const double zi = (lo + i < mid ? P.Q[(size_t)(mid - 1) * P.ldq + lo + i] : sgn * P.Q[(size_t)mid * P.ldq + lo + i]) * rs;
    dd[i] = di;
    z[i] = zi;
    mx = fmax(mx, fmax(fabs(di), fabs(zi)));
  }
  mx = block_max(mx, red);
  const double tol = 8.0 * EIGH_EPS * mx;
  // rank sort (ties by index)
  for (int i = tid; i < k; i += nt) {
    const double di = dd[i];
    int rk = 0;
    for (int u = 0; u < k; ++u) {
      const double du = dd[u];
      rk += (du < di || (du == di && u < i)) ? 1 : 0;
    }
    order[rk] = i;
    sd[rk] = di;
    sz[rk] = z[i];
  }
  __syncthreads();
  // vanishing weights out (to the end of the new order), survivors compacted in ascending order
  const int per = (k + nt - 1) / nt;
  const int t0 = tid * per, t1 = t0 + per < k ? t0 + per : k;
  int cnt = 0;
  for (int t = t0; t < t1; ++t) cnt += (rho * fabs(sz[t]) <= tol) ? 0 : 1;
  int ns;
  int at = block_excl_scan(cnt, sw, &ns);
  for (int t = t0; t < t1; ++t) {
    if (rho * fabs(sz[t]) <= tol) outp[k - 1 - (t - at)] = t;  // at = survivors before t, so (t - at) vanishing weights precede t
    else { surv[at] = t; ++at; }
  }
  __syncthreads();
  const int nsmall = k - ns;
  // close-pole test on the original values, every neighbouring pair of survivors
  for (int s2 = tid; s2 < ns; s2 += nt) {
    int f = 0;
    if (s2 > 0) {
      const int a = surv[s2 - 1], b2 = surv[s2];
      const double c = sz[b2], sn = sz[a], tt = sd[b2] - sd[a];
      f = fabs(tt * c * sn) <= tol * (c * c + sn * sn) ? 1 : 0;
    }
    flag[s2] = f;
  }
  __syncthreads();
  if (tid == 0) {
    int nr = 0;
    for (int s2 = 1; s2 < ns; ++s2) {
      if (flag[s2] != 1) continue;
      int a = surv[s2 - 1];
      int sa = s2 - 1;
      for (; s2 < ns; ++s2) {  // a run: the rotated pole / weight at `a` meets the next survivor
        const int b2 = surv[s2];
        double c = sz[b2], sn = sz[a];
        const double tt = sd[b2] - sd[a];
        const double t2 = c * c + sn * sn;
        if (!(fabs(tt * c * sn) <= tol * t2)) break;
        const double tau = sqrt(t2);
        c /= tau;
        sn = -sn / tau;
        sz[b2] = tau;
        sz[a] = 0.0;
        P.rot[lo + nr++] = make_double4((double)order[a], (double)order[b2], c, sn);
        const double d1 = sd[a] * c * c + sd[b2] * sn * sn;
        sd[b2] = sd[a] * sn * sn + sd[b2] * c * c;
        sd[a] = d1;
        flag[sa] = 2;
        a = b2;
        sa = s2;
      }
      // the pair that ended the run was tested with a rotated left member; the next pair starts from original values again
      if (s2 < ns && flag[s2] == 1) flag[s2] = 0;
    }
    P.nrot[blockIdx.x] = nr;
    P.rho[blockIdx.x] = rho;
  }
  __syncthreads();
  // second compaction: survivors that were not rotated away stay (ascending), the rotated-away ones join the deflated end
  const int per2 = (ns + nt - 1) / nt;
  const int u0 = tid * per2 < ns ? tid * per2 : ns, u1 = u0 + per2 < ns ? u0 + per2 : ns;
  int cnt2 = 0;
  for (int u = u0; u < u1; ++u) cnt2 += flag[u] == 2 ? 0 : 1;
  int K;
  int at2 = block_excl_scan(cnt2, sw, &K);
  for (int u = u0; u < u1; ++u) {
    if (flag[u] == 2) outp[k - 1 - nsmall - (u - at2)] = surv[u];
    else { outp[at2] = surv[u]; ++at2; }
  }
  if (tid == 0) P.Kp[blockIdx.x] = K;
  __syncthreads();
  for (int i = tid; i < k; i += nt) {
    const int t = outp[i];
    P.perm[lo + i] = order[t];
    P.lam[lo + i] = sd[t];
    if (i < K) { P.dl[lo + i] = sd[t]; P.wz[lo + i] = sz[t]; }
  }
}

// QB block = (QA block, rotated) with permuted columns. One CTA per row of a merge (grid.x = rows of the level, via row2merge).
__global__ void __launch_bounds__(256) dc_permute_kernel(const EighMerge* __restrict__ merges, const int* __restrict__ row2merge, const int* __restrict__ rows,
                                                         const double* __restrict__ QA, double* __restrict__ QB, int64_t ldq, const int* __restrict__ perm,
                                                         const double4* __restrict__ rot, const int* __restrict__ nrot) {
  extern __shared__ double rowv[];
  const int i = rows[blockIdx.x];
  const int mi = row2merge[blockIdx.x];
  const EighMerge mg = merges[mi];
  const int lo = mg.lo, k = mg.hi - mg.lo, tid = threadIdx.x;
  const double* src = QA + (size_t)i * ldq + lo;
  for (int c = tid; c < k; c += blockDim.x) rowv[c] = src[c];
  __syncthreads();
  const int nr = nrot[mi];
  if (tid == 0) {
    for (int t = 0; t < nr; ++t) {
      const double4 r = rot[lo + t];
      const int a = (int)r.x, b2 = (int)r.y;
      const double x = rowv[a], y = rowv[b2];
      rowv[a] = r.z * x + r.w * y;
      rowv[b2] = r.z * y - r.w * x;
    }
  }
  __syncthreads();
  double* dst = QB + (size_t)i * ldq + lo;
  for (int c = tid; c < k; c += blockDim.x) dst[c] = rowv[perm[lo + c]];
}

// secular equation: one warp per new column j of a merge. j < K: root in (dl_j, dl_{j+1}) as (origin, mu); row j of VT <- dl_i - lam_j
// (i < K), zeros beyond; j >= K (deflated): row j of VT <- e_j. The root finder is a bracketed rational iteration (below). NCH > 0: the merge has at most 32 NCH columns and each lane keeps its
// poles (relative to the origin) and weights in registers during the bisection; NCH = 0: any size, re-read from memory.
template <int NCH>
__global__ void __launch_bounds__(256) dc_secular_kernel(const EighMerge* __restrict__ merges, const int* __restrict__ Kp, const double* __restrict__ rhos,
                                                         const double* __restrict__ dl, const double* __restrict__ wz, double* __restrict__ lam,
                                                         double* __restrict__ VT, int64_t ldq) {
  const EighMerge mg = merges[blockIdx.y];
  const int lo = mg.lo, k = mg.hi - mg.lo;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int j = blockIdx.x * nw + wid;
  if (j >= k) return;
  const int K = Kp[blockIdx.y];
  double* vrow = VT + (size_t)(lo + j) * ldq + lo;
  if (j >= K) {
    for (int i = lane; i < k; i += 32) vrow[i] = i == j ? 1.0 : 0.0;
    return;
  }
  const double rho = rhos[blockIdx.y];
  const double* D = dl + lo;
  const double* W = wz + lo;
  const bool last = (j == K - 1);
  const double dj = D[j];
  const double gap = last ? 0.0 : D[j + 1] - dj;
  constexpr int NR = NCH > 0 ? NCH : 1;
  double tr[NR], rr[NR];  // pole - origin, rho w^2 (an absent pole: weight 0, far away)
  if (NCH > 0) {
#pragma unroll
    for (int m = 0; m < NR; ++m) {
      const int i = lane + 32 * m;
      const double wi = i < K ? W[i] : 0.0;
      rr[m] = rho * wi * wi;
      tr[m] = i < K ? D[i] - dj : 1e300;
    }
  }
  bool left = true;
  double hi = 0.0;
  if (!last) {
    const double half = 0.5 * gap;
    double f = 0.0;
    if (NCH > 0) {
#pragma unroll
      for (int m = 0; m < NR; ++m) f += rr[m] / (tr[m] - half);
    } else {
      for (int i = lane; i < K; i += 32) { const double wi = W[i]; f += rho * wi * wi / ((D[i] - dj) - half); }
    }
    f = 1.0 + warp_sum(f);
    left = f > 0.0;
    hi = half;
  } else {
    double s2 = 0.0;
    if (NCH > 0) {
#pragma unroll
      for (int m = 0; m < NR; ++m) s2 += rr[m];
    } else {
      for (int i = lane; i < K; i += 32) { const double wi = W[i]; s2 += rho * wi * wi; }
    }
    hi = warp_sum(s2);
  }
  const int o = left ? j : j + 1;
  const double dorg = D[o];
  if (NCH > 0 && !left) {
#pragma unroll
    for (int m = 0; m < NR; ++m) {
      const int i = lane + 32 * m;
      tr[m] = i < K ? D[i] - dorg : 1e300;
    }
  }
  const double sg = left ? 1.0 : -1.0;
  // nu = |mu| in (0, hi]. The bracket lives on the BIT PATTERNS of nu (lob: g still has the sign it has next to the origin pole, hib:
  // it has changed or vanished), so a bisection step halves the number of doubles left whatever their scale: <= 63 steps always
  // terminate. Most steps are not bisections: with g and g' at nu, the model c -/+ s / nu (the origin pole exact, everything else
  // constant) gives the next nu -- exact for one pole, quadratically convergent in general; a proposal that leaves the bracket or
  // a run of 40 steps falls back to bisection. Converged as LAPACK's dlaed4 judges it: |g| <= 8 eps (1 + sum |terms|).
  long long lob = 0, hib = __double_as_longlong(hi);
  double nu = hi;
  for (int it = 0; it < 112; ++it) {
    const double mu_t = sg * nu;
    double g = 0.0, gp = 0.0, ga = 0.0;
    if (NCH > 0) {
#pragma unroll
      for (int m = 0; m < NR; ++m) {
        const double q = 1.0 / (tr[m] - mu_t), t = rr[m] * q;
        g += t;
        ga += fabs(t);
        gp = fma(t, q, gp);
      }
    } else {
      for (int i = lane; i < K; i += 32) {
        const double wi = W[i];
        const double q = 1.0 / ((D[i] - dorg) - mu_t), t = rho * wi * wi * q;
        g += t;
        ga += fabs(t);
        gp = fma(t, q, gp);
      }
    }
    g = 1.0 + warp_sum(g);
    ga = 1.0 + warp_sum(ga);
    gp = warp_sum(gp);
    const long long bnu = __double_as_longlong(nu);
    const bool before = left ? (g < 0.0) : (g > 0.0);
    if (before) lob = bnu > lob ? bnu : lob; else hib = bnu < hib ? bnu : hib;
    if (fabs(g) <= 8.0 * EIGH_EPS * ga && ga < INFINITY) break;
    if (hib - lob <= 1) { nu = __longlong_as_double(hib); break; }
    const double den = left ? fma(gp, nu, g) : fma(gp, nu, -g);
    const double prop = gp * nu * nu / den;
    long long bn = (den > 0.0 && prop > 0.0 && prop < INFINITY) ? __double_as_longlong(prop) : -1;
    if (it >= 40 || !(bn > lob && bn < hib)) bn = lob + ((hib - lob) >> 1);
    nu = __longlong_as_double(bn);
  }
  const double mu = sg * nu;
  if (lane == 0) lam[lo + j] = dorg + mu;
  for (int i = lane; i < k; i += 32) vrow[i] = i < K ? (D[i] - dorg) - mu : 0.0;
}

// Loewner: what_i = sign(w_i) sqrt(-prod_j delta_j[i] / prod_{j != i} (dl_i - dl_j)). A CTA takes 32 consecutive i (lane = i: the reads of
// delta are coalesced), its 8 warps split the product over j and multiply their parts through shared memory.
__global__ void __launch_bounds__(256) dc_lowner_kernel(const EighMerge* __restrict__ merges, const int* __restrict__ Kp, const double* __restrict__ dl,
                                                        const double* __restrict__ wz, const double* __restrict__ VT, int64_t ldq, double* __restrict__ what) {
  __shared__ double part[8][32];
  const EighMerge mg = merges[blockIdx.y];
  const int lo = mg.lo;
  const int K = Kp[blockIdx.y];
  if (blockIdx.x * 32 >= K) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const double* D = dl + lo;
  double p = 1.0;
  if (i < K) {
    const double di = D[i];
    const double* col = VT + (size_t)lo * ldq + lo + i;
    for (int jb = wid; jb < K; jb += 8 * 4) {
      double num[4], den[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = jb + 8 * u;
        num[u] = j < K ? col[(size_t)j * ldq] : 1.0;
        den[u] = (j < K && j != i) ? di - D[j] : 1.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) p *= num[u] / den[u];
    }
  }
  part[wid][lane] = p;
  __syncthreads();
  if (wid == 0 && i < K) {
    double t = part[0][lane];
#pragma unroll
    for (int q = 1; q < 8; ++q) t *= part[q][lane];
    what[lo + i] = copysign(sqrt(fabs(t)), wz[lo + i]);
  }
}

// eigenvectors of the rank-one update: row j of VT <- normalised what_i / delta_j[i]; one warp per j < K
__global__ void __launch_bounds__(256) dc_vectors_kernel(const EighMerge* __restrict__ merges, const int* __restrict__ Kp, const double* __restrict__ what,
                                                         double* __restrict__ VT, int64_t ldq) {
  const EighMerge mg = merges[blockIdx.y];
  const int lo = mg.lo;
  const int K = Kp[blockIdx.y];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int j = blockIdx.x * nw + wid;
  if (j >= K) return;
  double* vrow = VT + (size_t)(lo + j) * ldq + lo;
  const double* Wh = what + lo;
  double s2 = 0.0;
  for (int i = lane; i < K; i += 32) {
    const double x = Wh[i] / vrow[i];
    vrow[i] = x;
    s2 += x * x;
  }
  s2 = warp_sum(s2);
  const double inv = 1.0 / sqrt(s2);
  for (int i = lane; i < K; i += 32) vrow[i] *= inv;
}

// Q block <- (permuted Q block) V for the SMALL merges of a level (k <= 128), all of them in one launch: C[i][j] = sum_l QB[i][l] VT[j][l].
// One CTA per 32 x 32 tile of one merge, operands staged in shared memory (odd pitch), four results per thread. The large merges go to
// the DMMA Gram engine, one launch each; for k <= 128 its launch costs more than the product (31 launches -> 9 at n = 1000).
constexpr int EIGH_SMALL_GEMM_MAX = 128;
__global__ void __launch_bounds__(256) dc_small_gemm_kernel(const EighMerge* __restrict__ merges, const double* __restrict__ QB, const double* __restrict__ VT,
                                                            double* __restrict__ QA, int64_t ldq) {
  extern __shared__ double gsm[];
  const EighMerge mg = merges[blockIdx.y];
  const int lo = mg.lo, k = mg.hi - mg.lo;
  const int T = (k + 31) >> 5;
  if ((int)blockIdx.x >= T * T) return;
  const int i0 = ((int)blockIdx.x / T) * 32, j0 = ((int)blockIdx.x % T) * 32;
  const int pitch = k | 1;
  double* As = gsm;
  double* Bs = gsm + 32 * pitch;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  for (int rr = ty; rr < 32; rr += 8) {
    const bool ia = i0 + rr < k, jb = j0 + rr < k;
    const double* qa = QB + (size_t)(lo + i0 + rr) * ldq + lo;
    const double* vb = VT + (size_t)(lo + j0 + rr) * ldq + lo;
    for (int l = tx; l < k; l += 32) {
      As[rr * pitch + l] = ia ? qa[l] : 0.0;
      Bs[rr * pitch + l] = jb ? vb[l] : 0.0;
    }
  }
  __syncthreads();
  double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
  const double* bp = Bs + tx * pitch;
  const double* ap = As + ty * pitch;
#pragma unroll 4
  for (int l = 0; l < k; ++l) {
    const double bv = bp[l];
    c0 = fma(ap[l], bv, c0);
    c1 = fma(ap[8 * pitch + l], bv, c1);
    c2 = fma(ap[16 * pitch + l], bv, c2);
    c3 = fma(ap[24 * pitch + l], bv, c3);
  }
  if (j0 + tx < k) {
    double* out = QA + (size_t)(lo + i0) * ldq + lo + j0 + tx;
    if (i0 + ty < k) out[(size_t)ty * ldq] = c0;
    if (i0 + ty + 8 < k) out[(size_t)(ty + 8) * ldq] = c1;
    if (i0 + ty + 16 < k) out[(size_t)(ty + 16) * ldq] = c2;
    if (i0 + ty + 24 < k) out[(size_t)(ty + 24) * ldq] = c3;
  }
}

// final order: rank of every eigenvalue (ascending, ties by index) -> order[rank] = index; W[rank] = lam * scales
__global__ void eigh_rank_kernel(const double* __restrict__ lam, int n, const double* __restrict__ scal, int* __restrict__ order, double* __restrict__ W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double li = lam[i];
  int rk = 0;
  for (int u = 0; u < n; ++u) {
    const double lu = lam[u];
    rk += (lu < li || (lu == li && u < i)) ? 1 : 0;
  }
  order[rk] = i;
  W[rk] = li * scal[1] * scal[0];
}

// out[k][i] = Q[i][order[k]]  (32 x 32 tiles through shared memory)
__global__ void eigh_gather_transpose_kernel(const double* __restrict__ Q, int64_t ldq, int n, const int* __restrict__ order, double* __restrict__ out,
                                             int64_t ldo) {
  __shared__ double t[32][33];
  const int k0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int kk = k0 + tx;
  const int col = kk < n ? order[kk] : 0;
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r;
    t[r][tx] = (i < n && kk < n) ? Q[(size_t)i * ldq + col] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, i = i0 + tx;
    if (k < n && i < n) out[(size_t)k * ldo + i] = t[tx][r];
  }
}

// rows of Z (eigenvectors of T, one per row) <- Q z with Q = H_0 H_1 ... H_{n-3}: reflectors applied last to first. One warp per
// vector, the vector in shared memory, reflector j read from row j of S (columns > j, v[j+1] = 1 stored explicitly).
__global__ void __launch_bounds__(256) eigh_backtransform_kernel(const double* __restrict__ S, int64_t lds, const double* __restrict__ tau, int n,
                                                                 double* __restrict__ Z, int64_t ldz) {
  extern __shared__ double zs[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int k = blockIdx.x * nw + wid;
  if (k >= n) return;
  double* z = zs + (size_t)wid * n;
  double* zrow = Z + (size_t)k * ldz;
  for (int c = lane; c < n; c += 32) z[c] = zrow[c];
  __syncwarp();
  for (int j = n - 3; j >= 0; --j) {
    const double tj = tau[j];
    if (tj == 0.0) continue;
    const double* vj = S + (size_t)j * lds;
    const int c0 = (j + 1) & ~31;
    double dot = 0.0;
    for (int c = c0 + lane; c < n; c += 32)
      if (c > j) dot += __ldg(vj + c) * z[c];
    dot = warp_sum(dot) * tj;
    for (int c = c0 + lane; c < n; c += 32)
      if (c > j) z[c] -= dot * __ldg(vj + c);
    __syncwarp();
  }
  for (int c = lane; c < n; c += 32) zrow[c] = z[c];
}

// ---- blocked back-transformation (n <= 2048). Reflectors are taken R at a time, last block first: block b holds reflectors
// jlow(b) <= j <= jtop(b) = n-3 - R b and H_jlow ... H_jtop = I - V T V' with the upper-triangular T of LAPACK's dlarft (forward,
// columnwise). eigh_larft_kernel builds every T (one CTA per block); eigh_backtransform_wy_kernel keeps one eigenvector per warp in
// REGISTERS (lane l holds entries l, l+32, ...), stages V block-wise in shared memory with cp.async (zero-filled up to the diagonal and
// beyond n, double buffered) and applies z <- z - V (T (V' z)): R independent dot products and R independent updates per stage instead
// of R dependent reflector applications.
__device__ __forceinline__ void cp_async8_zfill(uint32_t dst, const void* src, bool take) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(take ? 8 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int R>
__global__ void __launch_bounds__(256) eigh_larft_kernel(const double* __restrict__ S, int64_t lds, const double* __restrict__ tau, int n,
                                                         double* __restrict__ Tf) {
  __shared__ double G[R][R];   // G[a][b] = v_a . v_b (a < b), index = j - jlow
  __shared__ double T[R][R];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int jtop = n - 3 - R * (int)blockIdx.x;
  const int jlow = jtop - R + 1 > 0 ? jtop - R + 1 : 0;
  const int nb = jtop - jlow + 1;
  for (int pr = wid; pr < R * R; pr += 8) {
    const int a = pr / R, b = pr % R;
    if (a < b && b < nb) {
      // v_a has support > jlow + a, v_b > jlow + b >= the former
      const int jb = jlow + b;
      const double* va = S + (size_t)(jlow + a) * lds;
      const double* vb = S + (size_t)jb * lds;
      double acc = 0.0;
      for (int c = jb + 1 + lane; c < n; c += 32) acc += __ldg(va + c) * __ldg(vb + c);
      acc = warp_sum(acc);
      if (lane == 0) G[a][b] = acc;
    }
  }
  for (int t = tid; t < R * R; t += 256) T[t / R][t % R] = 0.0;
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < nb; ++i) {
      const double ti = tau[jlow + i];
      T[i][i] = ti;
      // T(0:i, i) = -tau_i T(0:i, 0:i) (V(:, 0:i)' v_i)
      for (int a = 0; a < i; ++a) {
        double acc = 0.0;
        for (int b = a; b < i; ++b) acc += T[a][b] * G[b][i];
        T[a][i] = -ti * acc;
      }
    }
  }
  __syncthreads();
  for (int t = tid; t < R * R; t += 256) Tf[(size_t)blockIdx.x * R * R + t] = T[t / R][t % R];
}

// NV eigenvectors per warp (8 per CTA of 8 / NV warps). The kernel reads one element of V from shared memory per FMA (0.62 ms at
// n = 1000 with NV = 1); NV = 2 halves that but leaves four warps per SM with 226 registers each and measured 0.66 ms: NV = 1 ships.
// The real fix is a DMMA form of the update (one LDS.128 per 512 FMAs), DESIGN.md section 8.
template <int NCH, int R, int NV>
__global__ void __launch_bounds__(256 / NV) eigh_backtransform_wy_kernel(const double* __restrict__ S, int64_t lds, const double* __restrict__ Tf, int n,
                                                                         double* __restrict__ Z, int64_t ldz) {
  extern __shared__ double vs[];  // 2 x R x (32 NCH)
  __shared__ double sT[2][R * R];
  constexpr int NP = 32 * NCH;
  constexpr int NT = 256 / NV;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int k0 = blockIdx.x * 8 + wid * NV;
  double z[NV][NCH];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const double* zrow = Z + (size_t)(k0 + v < n ? k0 + v : 0) * ldz;
#pragma unroll
    for (int m = 0; m < NCH; ++m) {
      const int c = lane + 32 * m;
      z[v][m] = (k0 + v < n && c < n) ? zrow[c] : 0.0;
    }
  }
  const uint32_t vs_u32 = (uint32_t)__cvta_generic_to_shared(vs);
  auto stage = [&](int blk, int jtop, int buf) {
    if (jtop >= 0) {
      const int jlow = jtop - R + 1 > 0 ? jtop - R + 1 : 0;
      const int cl = (jlow + 1) & ~31;
      for (int q = 0; q < R; ++q) {
        const int j = jlow + q;  // rows beyond the block (j > jtop) are zero-filled: they do nothing
        const double* src = S + (size_t)(j <= jtop ? j : jtop) * lds;
        const uint32_t dst = vs_u32 + (uint32_t)(((buf * R + q) * NP) * 8);
        for (int c = cl + tid; c < NP; c += NT) cp_async8_zfill(dst + c * 8, src + (c < n ? c : 0), j <= jtop && c > j && c < n);
      }
      for (int t = tid; t < R * R; t += NT) sT[buf][t] = Tf[(size_t)blk * R * R + t];
    }
    cp_async_commit();
  };
  stage(0, n - 3, 0);
  int buf = 0, blk = 0;
  for (int jtop = n - 3; jtop >= 0; jtop -= R, buf ^= 1, ++blk) {
    stage(blk + 1, jtop - R, buf ^ 1);  // the other buffer was consumed before the barrier that ended the previous round
    cp_async_wait<1>();
    __syncthreads();
    const int jlow = jtop - R + 1 > 0 ? jtop - R + 1 : 0;
    const int m0 = (jlow + 1) >> 5;
    const double* vb = vs + (size_t)(buf * R) * NP + lane;
    double y[NV][R];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int q = 0; q < R; ++q) y[v][q] = 0.0;
#pragma unroll
    for (int m = 0; m < NCH; ++m) {
      if (m >= m0) {
#pragma unroll
        for (int q = 0; q < R; ++q) {
          const double vq = vb[q * NP + 32 * m];
#pragma unroll
          for (int v = 0; v < NV; ++v) y[v][q] = fma(vq, z[v][m], y[v][q]);
        }
      }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int q = 0; q < R; ++q) y[v][q] = warp_sum(y[v][q]);
    double u[NV][R];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int a = 0; a < R; ++a) {
        double t = 0.0;
#pragma unroll
        for (int b = a; b < R; ++b) t += sT[buf][a * R + b] * y[v][b];
        u[v][a] = t;
      }
#pragma unroll
    for (int m = 0; m < NCH; ++m) {
      if (m >= m0) {
#pragma unroll
        for (int q = 0; q < R; ++q) {
          const double vq = vb[q * NP + 32 * m];
#pragma unroll
          for (int v = 0; v < NV; ++v) z[v][m] = fma(-vq, u[v][q], z[v][m]);
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if (k0 + v < n) {
      double* zrow = Z + (size_t)(k0 + v) * ldz;
#pragma unroll
      for (int m = 0; m < NCH; ++m) {
        const int c = lane + 32 * m;
        if (c < n) zrow[c] = z[v][m];
      }
    }
  }
}

// The same update on the FP64 TENSOR pipe (n <= 2048). The scalar kernel above reads one element of V from shared memory per FMA and
// is bound by exactly that (ncu: 0.62 ms at n = 1000, FP64 pipe 12 % busy); as DMMA.8x8x4 a 16-byte read feeds 512 FMAs.
// One CTA = 8 eigenvectors (the M of the MMA), its 8 warps own the 8-column tiles w, w+8, ... of them in registers, in the C-fragment
// layout (lane (g, t): Z[g][8 tile + 2t], [.. + 2t + 1]). Per block of 8 reflectors (V: 8 x n, zero outside the supports, staged with
// cp.async; T from eigh_larft_kernel<8>):
//   Y = Z V'   : the two C registers of a tile ARE the A fragments of two MMAs if the contraction runs over the columns {8 tile + 2t}
//                and {8 tile + 2t + 1} -- the order of a contraction is free -- and the matching B fragments V[g][8 tile + 2t (+1)]
//                are one LDS.128; partial Y of the 8 warps are summed through shared memory in a fixed order;
//   U = Y T'   : 12 FMAs per lane;
//   Z -= U V   : two MMAs per tile (reflectors 0-3 and 4-7), A = -U, B = V[t (+4)][8 tile + g], C = the tile.
template <int NT, bool DB>
__global__ void __launch_bounds__(256) eigh_backtransform_dmma_kernel(const double* __restrict__ S, int64_t lds, const double* __restrict__ Tf, int n,
                                                                      double* __restrict__ Z, int64_t ldz) {
  constexpr int R = 8;
  constexpr int NP = NT * 64;   // 8 warps x NT tiles x 8 columns
  constexpr int PV = NP + 8;    // pitch of a staged reflector: rows alternate between the two halves of the 128-byte bank line
  extern __shared__ double vsm[];  // (DB ? 2 : 1) x R x PV
  __shared__ double sT[2][R * R], sYp[8][64], sY[64];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  const int k0 = blockIdx.x * 8;
  const bool vrow = k0 + g < n;
  double* zrow = Z + (size_t)(vrow ? k0 + g : 0) * ldz;
  double z0[NT], z1[NT];
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int c = 8 * (w + 8 * i) + 2 * t;
    z0[i] = (vrow && c < n) ? zrow[c] : 0.0;
    z1[i] = (vrow && c + 1 < n) ? zrow[c + 1] : 0.0;
  }
  const uint32_t vs_u32 = (uint32_t)__cvta_generic_to_shared(vsm);
  auto stage = [&](int blk, int jtop, int buf) {
    if (jtop >= 0) {
      const int jlow = jtop - R + 1 > 0 ? jtop - R + 1 : 0;
      const int cl = (jlow + 1) & ~7;
      for (int q = 0; q < R; ++q) {
        const int j = jlow + q;  // rows beyond the block (j > jtop) are zero-filled: they do nothing
        const double* src = S + (size_t)(j <= jtop ? j : jtop) * lds;
        const uint32_t dst = vs_u32 + (uint32_t)(((buf * R + q) * PV) * 8);
        for (int c = cl + tid; c < NP; c += 256) cp_async8_zfill(dst + c * 8, src + (c < n ? c : 0), j <= jtop && c > j && c < n);
      }
      if (tid < R * R) sT[buf][tid] = Tf[(size_t)blk * R * R + tid];
    }
    cp_async_commit();
  };
  int buf = 0, blk = 0;
  if (DB) stage(0, n - 3, 0);
  for (int jtop = n - 3; jtop >= 0; jtop -= R, ++blk) {
    if (DB) {
      stage(blk + 1, jtop - R, buf ^ 1);  // the other buffer was consumed before the barrier that ended the previous round
      cp_async_wait<1>();
    } else {
      stage(blk, jtop, 0);
      cp_async_wait<0>();
    }
    __syncthreads();
    const int jlow = jtop - R + 1 > 0 ? jtop - R + 1 : 0;
    const int tile0 = (jlow + 1) >> 3;                 // first tile with a column inside a support
    const int i0 = tile0 > w ? (tile0 - w + 7) >> 3 : 0;  // first own tile >= tile0
    const double* vb = vsm + (size_t)(buf * R) * PV;
    // ---- Y = Z V' over the own tiles (four independent accumulator pairs)
    double ya[4] = {0.0, 0.0, 0.0, 0.0}, yb[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      if (i >= i0) {
        const double2 bv = *reinterpret_cast<const double2*>(vb + g * PV + 8 * (w + 8 * i) + 2 * t);
        dmma884(ya[i & 3], yb[i & 3], z0[i], bv.x);
        dmma884(ya[i & 3], yb[i & 3], z1[i], bv.y);
      }
    }
    sYp[w][g * 8 + 2 * t] = (ya[0] + ya[1]) + (ya[2] + ya[3]);
    sYp[w][g * 8 + 2 * t + 1] = (yb[0] + yb[1]) + (yb[2] + yb[3]);
    __syncthreads();
    if (tid < 64) {
      double acc = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) acc += sYp[q][tid];
      sY[tid] = acc;
    }
    __syncthreads();
    // ---- U = Y T' (upper-triangular T): this lane's A fragments of the update, negated
    double ua = 0.0, ub = 0.0;
#pragma unroll
    for (int b2 = 0; b2 < R; ++b2) {
      const double yv = sY[g * 8 + b2];
      if (b2 >= t) ua = fma(sT[buf][t * R + b2], yv, ua);
      if (b2 >= 4 + t) ub = fma(sT[buf][(4 + t) * R + b2], yv, ub);
    }
    ua = -ua;
    ub = -ub;
    // ---- Z -= U V
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      if (i >= i0) {
        const int c = 8 * (w + 8 * i) + g;
        const double b1 = vb[t * PV + c], b2 = vb[(4 + t) * PV + c];
        dmma884(z0[i], z1[i], ua, b1);
        dmma884(z0[i], z1[i], ub, b2);
      }
    }
    __syncthreads();
    if (DB) buf ^= 1;
  }
  cp_async_wait<0>();
  if (vrow) {
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const int c = 8 * (w + 8 * i) + 2 * t;
      if (c < n) zrow[c] = z0[i];
      if (c + 1 < n) zrow[c + 1] = z1[i];
    }
  }
}

}  // namespace plb
