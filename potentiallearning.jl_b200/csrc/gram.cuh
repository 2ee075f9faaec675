// gram.cuh — kernel-matrix (Gram) kernel: out[i][j] = epilogue( sum_k A[i][k] * B[j][k] ).
//
// Replaces the scalar pair loops of KernelMatrix (reference src/Kernels/kernels.jl:211-223 symmetric, :229-244
// rectangular) for compute_kernel(::DotProduct) (:30-36) and compute_kernel(::RBF) over Euclidean (:68-75,
// distances.jl:58-60). Both operands are K-major (row-major n x d, contraction index contiguous) — "NT" geometry.
//
// Design (B200 / sm_100a):
//   * persistent CTAs (one per SM), 128 x 128 output tile, BK = 16 doubles (= one 128-byte swizzle row) per stage,
//     PL_GRAM_STAGES-deep TMA -> shared-memory ring guarded by full/empty mbarriers;
//   * warps 8..11 form the TMA producer warpgroup (one elected lane works; the group hands its registers to the
//     consumers with setmaxnreg), warps 0..7 are DMMA consumers laid out 4 (m) x 2 (n),
//     each owning a 32 x 64 warp tile = 4 x 8 DMMA.8x8x4 tiles = 64 fp64 accumulators per thread;
//   * fragments are read with 128-bit LDS: one load yields the k-pair (2c, 2c+1) of one row, feeding two DMMAs
//     (the contraction index may be permuted freely as long as A and B agree). Bank conflicts are removed by a row
//     permutation inside each 8-row group (fragment row g <-> tile row perm(g) = ((g&1)<<2)|(g>>1)) so that every
//     quarter-warp touches 8 distinct 16-byte chunks of the 128B-swizzled tile; the permutation is undone in the
//     epilogue index math (free);
//   * symmetric mode computes only tiles tj >= ti, mirrors them on store (exact symmetry by construction), and on
//     diagonal tiles loads the operand tile once;
//   * the epilogue (normalise+power, or distance+exp) is fused; nothing but the final kernel matrix touches HBM.
#pragma once
#include "ptx.cuh"

namespace plb {

constexpr int GRAM_BM = 128;
constexpr int GRAM_BN = 128;
constexpr int GRAM_BK = 16;  // doubles per stage row: 128 bytes, the TMA 128B-swizzle span
constexpr int GRAM_STAGES = 5;
constexpr int GRAM_STAGES_TMAST = 4;            // operand ring of the TMA-store variant
constexpr int GRAM_STAGING_BYTES = 8 * 8192;    // HALF a 32 x 64 warp tile of staging per consumer warp (the epilogue leaves in two halves)
constexpr int GRAM_CONSUMER_WARPS = 8;
constexpr int GRAM_THREADS = (GRAM_CONSUMER_WARPS + 4) * 32;  // 2 DMMA warpgroups + 1 producer warpgroup (setmaxnreg needs whole warpgroups)
constexpr int PRODUCER_REGS = 24;
constexpr int CONSUMER_REGS = 240;  // per SMSP: 2 x 240 + 24 = 504 <= 512 registers per lane
constexpr int GRAM_TILE_BYTES = GRAM_BM * GRAM_BK * 8;  // 16 KB per operand per stage
constexpr int GRAM_SMEM_BYTES = GRAM_STAGES * 2 * GRAM_TILE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int GRAM_SMEM_BYTES_TMAST = GRAM_STAGES_TMAST * 2 * GRAM_TILE_BYTES + GRAM_STAGING_BYTES + 1024 + 256;
// RBF: table of 2^(j/256), j = 0..255, replicated 16 x (one copy per lane slot: conflict-free whatever the indices), kept in shared
// memory by every CTA of the exp epilogue (exp2_tab below)
constexpr int EXP_TAB_BITS = 8;
constexpr int EXP_TAB_N = 1 << EXP_TAB_BITS;
constexpr int EXP_TAB_BYTES = EXP_TAB_N * 16 * 8;
constexpr int GRAM_AUG_COLS = 2;  // RBF operands carry two extra contraction columns (u_i, 1) x (1, u_j): see rbf_aug_* in prelude.cuh

enum { EPI_LINEAR = 0, EPI_DOT = 1, EPI_RBF = 2, EPI_IMQ = 3 };

struct GramParams {
  int64_t n1, n2;       // output rows (A operand rows) / cols (B operand rows); symmetric: n2 == n1
  int64_t d;            // contraction length
  double* out;          // dense: row-major, leading dim ldo ; packed: tile-packed (128*128 doubles per tile)
  int64_t ldo;
  const double* ra;     // per-row statistics of the A operand rows (DOT: 1/sqrt(x.x); IMQ: quadratic form q_i; RBF: unused, see b_shift)
  const double* rb;     // same for B operand rows
  int symmetric;        // only tiles tj >= ti, mirrored per `fill`
  int same_operand;     // symmetric and A == B matrix: diagonal tiles load one operand tile
  int fill;             // PL_FILL_* bits (symmetric only)
  int packed;           // tile-packed output
  int alpha;            // DOT power
  double s;             // IMQ: 1 / ell^2 (RBF: folded into the operands)
  double beta;          // RBF scale ; IMQ: c^2
  double bmul;          // RBF: beta > 0 -> 1 (the scale is folded into the exponent by the prelude); else beta, multiplied in by the epilogue
  int T1, T2;           // tile rows / tile cols
  int64_t ntiles_total; // linear tile slots (symmetric: folded enumeration incl. a few invalid slots)
  int part, nparts;     // tile sharding across GPUs: slot l is ours iff l % nparts == part
  int tcol_begin;       // >= 0: "column range" enumeration (symmetric only): tiles (ti <= tj) of tile columns
  int tma_out;          // TMAST variant: the output matrix is described by the third tensor map (dense, 16-byte aligned, even ldo)
  int pdl;              // launched as a programmatic dependent of the prelude kernel: see pdl_wait() in the kernel
  int b_shift;          // RBF: the B operand reads its LAST k-step b_shift columns further right (the block that holds its variant of
                        //   the augmented columns); 0 otherwise
  const double* exp_tab;  // RBF: 2^(j/256) table in global memory, 256 doubles (replicated 16 x into shared memory by every CTA)
  int64_t tail_first;   // >= 0 (64 x 64 kernel only): TAIL mode. The last tail_count 128 x 128 tiles of this part (slots tail_first,
  int tail_count;       //   tail_first + nparts, ...) are computed as up to four 64 x 64 sub-tiles each, written into the parent tile's
  int T1p, T2p;         //   place (tile-packed output: inside the 128 x 128 packed tile). T1p / T2p: the 128-tile grid of the slots.
#ifdef PLB_GRAM_TRACE
  long long* trace;     // developer build (-DPLB_GRAM_TRACE, scripts/gram_trace.py): 8 int64 per CTA — globaltimer at entry, after the dependency
                        //   wait, after the first tile's main loop, after its epilogue, at the end of the tile loop (consumer warp 0), SM id, tiles done
#endif
  int tcol_end;         //       [tcol_begin, tcol_end), column by column — lets the host stream finished rows of the
                        //       Symmetric(:U) storage to the host while later columns are still being computed
};

// Folded enumeration of the upper triangle of a T x T tile grid: tile row p is paired with row T-1-p so that every
// pair has T+1 slots. Slot l -> (ti, tj); returns false for the (few) unused slots when T is odd.
__host__ __device__ __forceinline__ bool gram_slot_to_tile(const int T1, const int T2, const int symmetric, int64_t l, int& ti, int& tj) {
  if (!symmetric) {
    ti = (int)(l / T2);
    tj = (int)(l % T2);
    return ti < T1;
  }
  const int T = T1;
  const int p = (int)(l / (T + 1));
  const int q = (int)(l % (T + 1));
  if (q < T - p) {
    ti = p;
    tj = p + q;
    return true;
  }
  const int p2 = T - 1 - p;
  if (p2 <= p) return false;  // middle row of an odd T is not paired
  ti = p2;
  tj = p2 + (q - (T - p));
  return true;
}
// column-range enumeration: slot l (0-based inside the range) -> (ti, tj), tj in [c0, c1), ti <= tj, column-major over the
// upper triangle: global index L = c0(c0+1)/2 + l, tj = floor((sqrt(8L+1)-1)/2), ti = L - tj(tj+1)/2
__host__ __device__ __forceinline__ void gram_colrange_slot_to_tile(int c0, int64_t l, int& ti, int& tj) {
  const int64_t L = (int64_t)c0 * (c0 + 1) / 2 + l;
  int64_t t = (int64_t)((sqrt(8.0 * (double)L + 1.0) - 1.0) * 0.5);
  while (t * (t + 1) / 2 > L) --t;
  while ((t + 1) * (t + 2) / 2 <= L) ++t;
  tj = (int)t;
  ti = (int)(L - t * (t + 1) / 2);
}
__host__ __device__ __forceinline__ int64_t gram_colrange_num_slots(int c0, int c1) {
  return (int64_t)c1 * (c1 + 1) / 2 - (int64_t)c0 * (c0 + 1) / 2;
}
__host__ __device__ __forceinline__ int64_t gram_num_slots(int T1, int T2, int symmetric) {
  if (!symmetric) return (int64_t)T1 * T2;
  return (int64_t)((T1 + 1) / 2) * (T1 + 1);
}

// One unit of work of a CTA: the output tile (ti, tj) in units of the kernel's own tile edge, the packed tile it lands in (pk) and, in TAIL
// mode, its offset (oi, oj) inside that 128 x 128 parent and whether the parent is a diagonal tile (its (0, 1) sub-tile is mirrored into it).
struct GramTile {
  int ti, tj, oi, oj;
  int64_t pk;
  bool diag_parent;
};
template <int NARROW>
__device__ __forceinline__ bool gram_tile_of(const GramParams& p, int64_t l, GramTile& t) {
  t.oi = t.oj = 0;
  t.diag_parent = false;
  if (NARROW == 2 && p.tail_first >= 0) {
    const int64_t slot = p.tail_first + (l >> 2) * p.nparts;
    int pi, pj;
    if (!gram_slot_to_tile(p.T1p, p.T2p, p.symmetric, slot, pi, pj)) return false;
    const int a = (int)((l >> 1) & 1), b = (int)(l & 1);
    if (p.symmetric && pi == pj && a > b) return false;  // produced by mirroring sub-tile (0, 1): never computed independently
    t.ti = 2 * pi + a;
    t.tj = 2 * pj + b;
    if ((int64_t)t.ti * 64 >= p.n1 || (int64_t)t.tj * 64 >= p.n2) return false;
    t.pk = (slot - p.part) / p.nparts;
    t.oi = 64 * a;
    t.oj = 64 * b;
    t.diag_parent = p.symmetric && pi == pj;
    return true;
  }
  if (p.tcol_begin >= 0) gram_colrange_slot_to_tile(p.tcol_begin, l, t.ti, t.tj);
  else if (!gram_slot_to_tile(p.T1, p.T2, p.symmetric, l, t.ti, t.tj)) return false;
  t.pk = (l - p.part) / p.nparts;
  return true;
}

// Integer power by squaring for a general DotProduct alpha (kernels.jl:35 `^d.α`); alpha = 1, 2 are inlined in the epilogue.
__device__ __noinline__ double ipow_general(double x, int a) {
  const bool neg = a < 0;
  unsigned e = neg ? (unsigned)(-a) : (unsigned)a;
  double r = 1.0, b = x;
  while (e) {
    if (e & 1u) r *= b;
    b *= b;
    e >>= 1;
  }
  return neg ? 1.0 / r : r;
}

// 2^(y / 256) — the whole RBF epilogue. The accumulator of the RBF Gram kernel already IS
//   y = 256 log2(e) * (-0.5 d2 / ell^2 + ln beta)        (kernels.jl:73-74 with distances.jl:59 expanded)
// because its operands are pre-scaled by sqrt(256 log2(e)) / ell and carry the row terms -q_i/2 as two augmented contraction
// columns (u_i, 1) x (1, u_j) (prelude.cuh, rbf_aug_kernel): no per-element argument arithmetic is left. Then
//   N = rint(y) by the magic-number add, r = y - N (exact), j = N mod 256 -> T = 2^(j/256) from shared memory, n = N div 256,
//   2^(y/256) = 2^n * T * e^(c r),  c = ln2 / 256,  |c r| <= 1.36e-3:  e^z - 1 = z + z^2 (b0 + b1 z + b2 z^2), the b's interpolated at
//   Chebyshev nodes (max relative error 9.5e-18), 2^n applied to the exponent field.
// 8 FP64-pipe instructions per element (3 DADD, 4 DFMA, 1 DMUL) instead of the 22 of the Cody-Waite + degree-13 polynomial of round 1
// — scalar FP64 and DMMA share one pipe, so every one of them is time taken from the DMMAs. The table is replicated 16 x in shared
// memory (entry j of lane-slot s = lane & 15 at (16 j + s) * 8: 32 KB): every half-warp reads 16 distinct bank pairs whatever the
// indices are — a 4096-entry unreplicated table (degree 3, 7 instructions) measured 6.2 conflict wavefronts per lookup and made the
// epilogue LSU-bound (profiles/r02b notes). Max relative error 2.2e-16 (1 ulp: half an ulp of the table entry + the final
// rounding) against long double over 2e7 points (tests/test_fast_exp.py re-states it). Branch free; y > 0 (beta > 1) works the same
// way. Arguments below -1022 * 256 (x < -708.4: results under 2.3e-308, denormal or zero in the reference) come out as a number
// <= 4.5e-308 (N is clamped, below). |y| < 2^31 is assumed (|x| < 5.8e6: beyond that N is garbage, the masked table index stays
// in bounds, and the result is some number <= 4.5e-308 or, for y = -inf / NaN, NaN).
__constant__ double EXP2_C[5] = {6755399441055744.0 /*1.5 * 2^52*/, 0.0027076061740622863 /*c = ln2/256*/, 3.6655655969101062e-06 /*b0 c^2*/,
                                 3.308302907918888e-09 /*b1 c^3*/, 2.239395293483284e-12 /*b2 c^4*/};
// tab.addr = table base + 8 * (lane & 15); entry offset (N mod 256) * 128 (one LOP3 + one multiply-add).
// Instruction diet (the epilogue of the staged kernel is bound by ISSUE slots: 2 warps x 64 elements per sub-partition and tile): the flush is
// an integer max on N (N >= -1022 * 256 keeps the exponent field in range; everything below comes out as <= 2^-1021 = 4.5e-308 instead
// of exactly 0 — the reference has denormals or 0 there), the exponent goes in with one shift and one multiply-add.
constexpr int EXP2_NMIN = -1022 * EXP_TAB_N;
struct Exp2Tab {
  uint32_t addr;  // shared-memory address of the table + 8 * (lane & 15)
};
__device__ __forceinline__ double exp2_tab(const double y, const Exp2Tab tab) {
  const double t = y + EXP2_C[0];
  const int N = max(__double2loint(t), EXP2_NMIN);
  const double r = y - (t - EXP2_C[0]);
  // integer side spelled in PTX so that it stays 2 + 2 instructions (and / mad for the table address, shr / mad for the exponent field);
  // left to itself the compiler rewrites both into shift + and + add
  uint32_t a;
  int hi_add;
  asm("{\n"
      ".reg .b32 j;\n"
      "and.b32 j, %2, 255;\n"
      "mad.lo.u32 %0, j, 128, %3;\n"
      "shr.s32 j, %2, 8;\n"
      "mul.lo.s32 %1, j, 1048576;\n"
      "}\n"
      : "=r"(a), "=r"(hi_add)
      : "r"(N), "r"(tab.addr));
  double T;
  asm("ld.shared.f64 %0, [%1];" : "=d"(T) : "r"(a));
  double q = fma(r, EXP2_C[4], EXP2_C[3]);
  q = fma(q, r, EXP2_C[2]);
  q = fma(q, r, EXP2_C[1]);
  const double e = fma(T * r, q, T);
  return __hiloint2double(__double2hiint(e) + hi_add, __double2loint(e));
}
static_assert(EXP_TAB_BITS == 8, "exp2_tab spells 255 / 8 in its PTX");

// Fused epilogue of one warp tile. GENERAL = false: interior, off-diagonal tile (97% of the tiles of a large matrix):
// no bounds or triangle logic, ~30 instructions per element. GENERAL = true: edge and diagonal tiles.
//   rel0: (j - i) of accumulator (0,0,0) for diagonal tiles (elements with j < i are produced by mirroring, j == i is the
//         kernel diagonal), 4096 otherwise. w1 / w2: write the direct / mirrored position; in diagonal tiles both
//         positions are always written and hold 0 when their triangle was not requested (zeros(n,n) storage).
template <int EPI, bool GENERAL, int ALPHA /* 2: DotProduct power 2 / RBF scale folded; 0: general (runtime p.alpha, p.bmul) */, int MT, int NT>
__device__ __forceinline__ void gram_epilogue(const double (&acc)[MT][NT][2], const GramParams& p, int64_t i0, int64_t j0, int rel0, bool w1, bool w2,
                                              double* __restrict__ prow, double* __restrict__ pcol, int64_t ldo, const Exp2Tab tab) {
  // per-row / per-column statistics: DotProduct 1/|x|, IMQ the quadratic forms; RBF has them inside the accumulator (augmented columns)
  constexpr bool STATS = EPI == EPI_DOT || EPI == EPI_IMQ;
  unsigned imask = (1u << MT) - 1u;
  double sa[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    bool ok = true;
    if (GENERAL) {
      ok = i0 + 8 * mt < p.n1;
      if (!ok) imask &= ~(1u << mt);
    }
    const double r = (STATS && ok) ? __ldg(p.ra + i0 + 8 * mt) : 0.0;
    sa[mt] = EPI == EPI_IMQ ? p.s * r : r;
  }
  const double m2s = -2.0 * p.s;
  // k(x,x): beta for RBF (d2 == 0), c2^(-1/2) for IMQ, n/sqrt(n*n) = 1 for DotProduct
  const double diagval = (EPI == EPI_RBF) ? p.beta : (EPI == EPI_IMQ ? rsqrt(p.beta) : 1.0);
  const bool fill1 = p.packed || !p.symmetric || (p.fill & 1);
  const bool fill2 = p.packed || (p.fill & 2);
  const bool dg = rel0 != 4096;
  // 8 independent elements per step (one row, 4 column blocks x 2 columns = 2 full 128-byte lines of the direct store)
  // keep the FP64 pipe busy through the exp chains without blowing up the live register set
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    double* const pr = prow + (int64_t)(8 * mt) * ldo;
#pragma unroll
    for (int ng = 0; ng < NT / 4; ++ng) {
      double v[4][2];
      bool jok[4][2];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int nt = 4 * ng + q;
          jok[q][c] = GENERAL ? (j0 + 8 * nt + 4 * c < p.n2) : true;
          const double r = (STATS && jok[q][c]) ? __ldg(p.rb + j0 + 8 * nt + 4 * c) : 0.0;
          double x = acc[mt][nt][c];
          if (EPI == EPI_DOT) {
            x = x * sa[mt] * r;
            if (ALPHA == 2) x = x * x;
            else if (ALPHA == 0) x = ipow_general(x, p.alpha);
          } else if (EPI == EPI_RBF) {
            // the accumulator is 256 log2(e) (-0.5 d2 / ell^2 + ln beta)  (kernels.jl:73-74 with distances.jl:59 expanded)
            x = exp2_tab(x, tab);
            if (ALPHA == 0) x *= p.bmul;  // only for beta <= 0; for beta > 0 the scale is folded into the exponent (bmul == 1)
          } else if (EPI == EPI_IMQ) {
            // (c2 + d2/ell^2)^(-1/2)  (kernels.jl:156-164): d2/ell^2 = s*(q_i + q_j - 2 G_ij) with s = 1/ell^2
            x = rsqrt(p.beta + fma(x, m2s, fma(p.s, r, sa[mt])));
          }
          v[q][c] = x;
        }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int nt = 4 * ng + q;
          double x = v[q][c];
          if (!GENERAL) {
            if (w1) pr[8 * nt + 4 * c] = x;
            if (w2) pcol[(int64_t)(8 * nt + 4 * c) * ldo + 8 * mt] = x;
          } else {
            const bool ok = ((imask >> mt) & 1u) && jok[q][c];
            const int rel = rel0 + (8 * nt + 4 * c - 8 * mt);
            // k(x,x): beta (d2 == 0), resp. n/sqrt(n*n) = 1 in the reference — unless it is NaN there too (a zero feature vector gives
            // 0/sqrt(0) in DotProduct, NaN inputs propagate): x == x keeps those
            if (EPI != EPI_LINEAR && rel == 0 && x == x) x = diagval;
            if (ok && w1 && rel >= 0) pr[8 * nt + 4 * c] = (!dg || fill1 || rel == 0) ? x : 0.0;
            if (ok && w2 && rel > 0) pcol[(int64_t)(8 * nt + 4 * c) * ldo + 8 * mt] = (!dg || fill2) ? x : 0.0;
          }
        }
    }
  }
}

// Shared-memory staged, TMA-stored epilogue of one interior off-diagonal 32 x 64 warp tile (TMAST kernel variant).
// The fragment stores of gram_epilogue touch 8 rows x 32 B (direct) or 4 rows x 64 B (mirror) per instruction and the warp sits through
// them: with the 8-instruction exponential the register epilogue of the RBF kernel was bound by the LSU, not by the FP64 pipe (ncu,
// profiles/r02b notes); shallow contractions (d <= 128: config 1's d = 26) are bound by WRITING K outright. Here the values go to an
// 8 KB per-warp staging buffer in the 128B-swizzled box layout (conflict-free st.shared), one lane hands 16 x 32 boxes to the TMA unit
// (full 128-byte lines, asynchronous) and the warp moves on to the next tile while the writes drain. The warp tile leaves in two
// halves through the same buffer (direct copy: column halves; mirrored copy, staged transposed: row halves), the values of the second
// half being computed while the TMA unit reads the first — half the staging memory buys a fourth operand stage and the RBF table.
//   direct half h (columns 32h .. 32h+31, 32 rows), box b = 0, 1 (16 columns each): stg + 4096 b + 128 row + ((8 (col & 15)) ^ (16 (row & 7)))
//   mirror half ih (i-columns 16 ih .. +15), box jh = 0, 1 (32 j-rows each):          stg + 4096 jh + 128 (j & 31) + ((8 (i & 15)) ^ (16 (j & 7)))
template <int EPI>
__device__ __forceinline__ void gram_epilogue_staged(double (&acc)[4][8][2], const GramParams& p, const int64_t i0, const int64_t j0, const bool f1,
                                                     const bool f2, const uint32_t stg, const CUtensorMap* tmO, const int gi0w, const int gj0w,
                                                     const int lane, const int pg, const int t, const Exp2Tab tab) {
  constexpr bool STATS = EPI == EPI_DOT || EPI == EPI_IMQ;
  const double m2s = -2.0 * p.s;
  // values in place (the common parameters only: DotProduct alpha = 2, RBF scale folded into the exponent), rows [8 m0, 8 m1)
  auto values = [&](const int m0, const int m1) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      if (mt < m0 || mt >= m1) continue;
      const double ri = STATS ? __ldg(p.ra + i0 + 8 * mt) : 0.0;
      const double sa = EPI == EPI_IMQ ? p.s * ri : ri;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const double r = STATS ? __ldg(p.rb + j0 + 8 * nt + 4 * c) : 0.0;
          double x = acc[mt][nt][c];
          if (EPI == EPI_DOT) {
            x = x * sa * r;
            x = x * x;
          } else if (EPI == EPI_RBF) {
            x = exp2_tab(x, tab);
          } else if (EPI == EPI_IMQ) {
            x = rsqrt(p.beta + fma(x, m2s, fma(p.s, r, sa)));
          }
          acc[mt][nt][c] = x;
        }
    }
  };
  // the staging buffer is free once the bulk stores issued from it have finished READING it
  auto buffer_free = [&]() {
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
  };
  if (f1) {
    values(0, 4);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      buffer_free();
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int nt = 4 * h + q;
            const int col = 8 * nt + t + 4 * c;  // row = 8 mt + pg, row & 7 = pg
            sts64(stg + (uint32_t)(((col >> 4) & 1) * 4096 + (8 * mt + pg) * 128 + (((col & 15) << 3) ^ (pg << 4))), acc[mt][nt][c]);
          }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int b = 0; b < 2; ++b) tma_store_2d(tmO, stg + b * 4096, gj0w + 32 * h + 16 * b, gi0w);
        tma_store_commit();
      }
      __syncwarp();
    }
  }
  if (f2) {
#pragma unroll
    for (int ih = 0; ih < 2; ++ih) {
      if (!f1) values(2 * ih, 2 * ih + 2);  // the second half's values are computed while the TMA unit reads the first
      buffer_free();
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int mt = 2 * ih + m;
            const int il = 8 * mt + pg, jl = 8 * nt + t + 4 * c;  // jl & 7 = t + 4 c
            sts64(stg + (uint32_t)((nt >> 2) * 4096 + (jl & 31) * 128 + (((il & 15) << 3) ^ ((t + 4 * c) << 4))), acc[mt][nt][c]);
          }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) tma_store_2d(tmO, stg + jh * 4096, gi0w + 16 * ih, gj0w + 32 * jh);
        tma_store_commit();
      }
      __syncwarp();
    }
  }
}

// NARROW is the tile SHAPE:
// NARROW = 0: 128 x 128 output tile, warps 4 (m) x 2 (n), warp tile 32 x 64.
// NARROW = 2: 64 x 64 output tile, warps 4 (m) x 2 (n), warp tile 16 x 32: small matrices whose 128 x 128 tiles would not fill the machine
//             (config 1, the kDPP example shape N = 1000: 36 tiles of 128 on 148 SMs, 136 tiles of 64).
// NARROW = 1: 128 x 32 output tile for rectangular results with few columns (projections W'(x - m) of transform!, pca_state.jl:42-62;
//             KernelMatrix(F1, F2, k) against a handful of reference points): warps 8 (m) x 1, warp tile 16 x 32. A 128-wide tile
//             would spend 128 / n2 times the DMMAs such a result needs; at n2 <= 32 the narrow tile is within 1.4 x of the HBM time of
//             streaming the tall operand once.
// TMAST = 1 (NARROW = 0 only): 4-stage operand ring + 8 x 8 KB staging buffers; interior off-diagonal tiles of a dense result leave
//             through gram_epilogue_staged (tmO describes the output matrix with a 16 x 32 box).
template <int EPI, int NARROW, int TMAST>
__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO, const GramParams p) {
  constexpr int NSTAGE = TMAST ? GRAM_STAGES_TMAST : GRAM_STAGES;
  constexpr int BM = NARROW == 2 ? 64 : GRAM_BM;
  constexpr int BN = NARROW == 1 ? 32 : (NARROW == 2 ? 64 : GRAM_BN);
  constexpr int MT = NARROW ? 2 : 4;   // 8-row DMMA m-tiles per warp
  constexpr int NT = NARROW ? 4 : 8;   // 8-column DMMA n-tiles per warp
  constexpr uint32_t A_TILE_BYTES = BM * GRAM_BK * 8;
  constexpr uint32_t B_TILE_BYTES = BN * GRAM_BK * 8;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned tiles
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t tab_base = smem_base + NSTAGE * 2 * GRAM_TILE_BYTES + (TMAST ? GRAM_STAGING_BYTES : 0);  // RBF: 2^(j/256) table, 16 replicas
  const uint32_t bar_base = tab_base + (EPI == EPI_RBF ? EXP_TAB_BYTES : 0);  // full[S] then empty[S], 8 bytes each
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (NSTAGE + s); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), GRAM_CONSUMER_WARPS);
    }
    fence_barrier_init();
  }
  if (EPI == EPI_RBF) {  // the table is written once at pl_init: no dependence on the prelude kernel
    for (int e = threadIdx.x; e < EXP_TAB_N * 8; e += GRAM_THREADS) {  // 128-bit store e covers lane slots 2 (e & 7), +1 of entry e >> 3
      const double v = __ldg(p.exp_tab + (e >> 3));
      asm volatile("st.shared.v2.f64 [%0], {%1, %1};" ::"r"(tab_base + 16u * e), "d"(v) : "memory");
    }
  }
  __syncthreads();
  // Launched as a programmatic dependent of the prelude kernel (row statistics / centred copy): everything above overlaps its tail.
  // DotProduct contracts the caller's X itself, so only the epilogue (row statistics) has to wait; the others read the prelude's output.
#ifdef PLB_GRAM_TRACE
  auto gtime = []() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
  long long* const tr = p.trace ? p.trace + 8 * (int64_t)blockIdx.x : nullptr;
  if (tr && threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tr[0] = gtime();
    tr[5] = smid;
  }
#endif
  if (p.pdl && EPI != EPI_DOT) pdl_wait();
#ifdef PLB_GRAM_TRACE
  if (tr && threadIdx.x == 0) tr[1] = gtime();
  int tiles_done = 0;
#endif

  const int KT = (int)((p.d + GRAM_BK - 1) / GRAM_BK);
  const bool tail = NARROW == 2 && p.tail_first >= 0;
  const int64_t slot0 = tail ? (int64_t)blockIdx.x : (int64_t)blockIdx.x * p.nparts + p.part;
  const int64_t slot_stride = tail ? (int64_t)gridDim.x : (int64_t)gridDim.x * p.nparts;
  const int64_t slot_end = tail ? (int64_t)4 * p.tail_count : p.ntiles_total;

  if (warp >= GRAM_CONSUMER_WARPS) {
    // ------------------------------ TMA producer warpgroup (one elected lane works) ------------------------------
    setmaxnreg_dec<PRODUCER_REGS>();
    if (warp == GRAM_CONSUMER_WARPS && lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      uint32_t it = 0;
      for (int64_t l = slot0; l < slot_end; l += slot_stride) {
        GramTile tl;
        if (!gram_tile_of<NARROW>(p, l, tl)) continue;
        const int ti = tl.ti, tj = tl.tj;
        const bool one_tile = p.same_operand && (ti == tj);
        const uint32_t bytes = one_tile ? A_TILE_BYTES : A_TILE_BYTES + B_TILE_BYTES;
        for (int kt = 0; kt < KT; ++kt, ++it) {
          const int s = it % NSTAGE;
          const uint32_t ph = (it / NSTAGE) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_arrive_expect_tx(full_bar(s), bytes);
          const uint32_t dstA = smem_base + s * 2 * GRAM_TILE_BYTES;
          tma_load_2d(dstA, &tmA, full_bar(s), kt * GRAM_BK, ti * BM);
          if (!one_tile) tma_load_2d(dstA + GRAM_TILE_BYTES, &tmB, full_bar(s), kt * GRAM_BK + (kt == KT - 1 ? p.b_shift : 0), tj * BN);
        }
      }
    }
    return;
  }

  // ------------------------------ DMMA consumers ------------------------------
  setmaxnreg_inc<CONSUMER_REGS>();
  const int wm = NARROW == 1 ? warp : warp >> 1;  // 8*MT-row slab
  const int wn = NARROW == 1 ? 0 : warp & 1;      // 8*NT-col slab
  const int g = lane >> 2;
  const int t = lane & 3;
  const int pg = ((g & 1) << 2) | (g >> 1);  // tile row (within an 8-row group) served by fragment row g
  // byte offset of this thread's 16B chunk inside a 128B row for k-half 0; k-half 1 is the same ^ 64
  const uint32_t xo = (uint32_t)((t ^ pg) << 4);
  const uint32_t offA0 = (uint32_t)((8 * MT * wm + pg) * 128) + xo;  // + mt * 1024
  const uint32_t offB0 = (uint32_t)((8 * NT * wn + pg) * 128) + xo;  // + nt * 1024

  uint32_t it = 0;
  for (int64_t l = slot0; l < slot_end; l += slot_stride) {
    GramTile tl;
    if (!gram_tile_of<NARROW>(p, l, tl)) continue;
    const int ti = tl.ti, tj = tl.tj;
    const bool one_tile = p.same_operand && (ti == tj);

    double acc[MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    // Software-pipelined main loop: the fragments of the next half-step (and, across k-iterations, the barrier wait of the
    // next stage) are issued before the 64 DMMAs of the current half-step, so LDS latency and mbarrier wake-up never
    // expose the DMMA pipe. Two fragment buffers of 12 x 128 bit (96 registers) on top of the 128 accumulator registers.
    double2 a0[MT], b0[NT], a1[MT], b1[NT];
    {
      const int s = it % NSTAGE;
      mbar_wait(full_bar(s), (it / NSTAGE) & 1u);
      const uint32_t sA = smem_base + s * 2 * GRAM_TILE_BYTES;
      const uint32_t sB = one_tile ? sA : sA + GRAM_TILE_BYTES;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) a0[mt] = lds128(sA + (offA0 + mt * 1024));
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) b0[nt] = lds128(sB + (offB0 + nt * 1024));
    }
    for (int kt = 0; kt < KT; ++kt, ++it) {
      const int s = it % NSTAGE;
      const uint32_t sA = smem_base + s * 2 * GRAM_TILE_BYTES;
      const uint32_t sB = one_tile ? sA : sA + GRAM_TILE_BYTES;
      // second k-half of this stage -> buffer 1, then the stage can go back to the producer
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) a1[mt] = lds128(sA + ((offA0 + mt * 1024) ^ 64u));
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) b1[nt] = lds128(sB + ((offB0 + nt * 1024) ^ 64u));
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a0[mt].x, b0[nt].x);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a0[mt].y, b0[nt].y);
      if (kt + 1 < KT) {
        const int s2 = (it + 1) % NSTAGE;
        mbar_wait(full_bar(s2), ((it + 1) / NSTAGE) & 1u);
        const uint32_t sA2 = smem_base + s2 * 2 * GRAM_TILE_BYTES;
        const uint32_t sB2 = one_tile ? sA2 : sA2 + GRAM_TILE_BYTES;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) a0[mt] = lds128(sA2 + (offA0 + mt * 1024));
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) b0[nt] = lds128(sB2 + (offB0 + nt * 1024));
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a1[mt].x, b1[nt].x);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a1[mt].y, b1[nt].y);
      // all fragments of stage s have been consumed from registers (a DMMA reads its operands at issue): release the stage
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(s));
    }

#ifdef PLB_GRAM_TRACE
    if (tr && threadIdx.x == 0 && tiles_done == 0) tr[2] = gtime();
#endif
    // ------------------------------ fused epilogue ------------------------------
    // accumulator (mt, nt, c) of lane (g, t) is element
    //   i = 128*ti + 32*wm + 8*mt + perm(g),   j = 128*tj + 64*wn + 8*nt + perm(2t + c) = ... + t + 4c
    const int il0 = 8 * MT * wm + pg;  // local row of mt = 0
    const int jl0 = 8 * NT * wn + t;   // local col of (nt = 0, c = 0)
    if (p.pdl && EPI == EPI_DOT) pdl_wait();  // the row statistics come from the prelude kernel this launch depends on
    const int64_t i0 = (int64_t)ti * BM + il0;
    const int64_t j0 = (int64_t)tj * BN + jl0;
    const bool diag_tile = p.symmetric && (ti == tj);
    const bool f1 = p.packed || !p.symmetric || (p.fill & 1);            // write [i][j]
    const bool f2 = !p.packed && p.symmetric && (p.fill & 2);            // write [j][i] (off-diagonal tiles)
    double* base;
    int64_t ldo, ib, jb;
    if (p.packed) {  // tile pk of this part, row-major; in TAIL mode the 64 x 64 sub-tile sits at (oi, oj) of its 128 x 128 parent
      base = p.out + tl.pk * (int64_t)(tail ? GRAM_BM * GRAM_BN : BM * BN);
      ldo = tail ? GRAM_BN : BN;
      ib = tl.oi + il0;
      jb = tl.oj + jl0;
    } else {
      base = p.out;
      ldo = p.ldo;
      ib = i0;
      jb = j0;
    }
    double* const prow = base + ib * ldo + jb;  // direct:  prow[8*mt*ldo + 8*nt + 4*c]
    double* const pcol = base + jb * ldo + ib;  // mirror:  pcol[(8*nt + 4*c)*ldo + 8*mt]
    const bool interior = ((int64_t)(ti + 1) * BM <= p.n1) && ((int64_t)(tj + 1) * BN <= p.n2);
    const int rel0 = diag_tile ? (jl0 - il0) : 4096;
    // TAIL mode, sub-tile (0, 1) of a DIAGONAL parent: its mirror image is sub-tile (1, 0) of the same packed tile
    const bool mirror_in_parent = tail && p.packed && tl.diag_parent && !diag_tile;
    const bool f2m = f2 || mirror_in_parent;
    const bool w1 = f1 || diag_tile, w2 = diag_tile ? true : f2m;
    // ALPHA doubles as the "uncommon parameters" switch: DotProduct power != 2, or an RBF scale beta <= 0 that cannot be
    // folded into the exponent: one general instantiation
    const bool common = (EPI == EPI_DOT) ? (p.alpha == 2) : (EPI == EPI_RBF ? (p.bmul == 1.0) : true);
    if (TMAST && !NARROW && common && interior && !diag_tile && p.tma_out) {
      if constexpr (TMAST && !NARROW) {
        const uint32_t stg = smem_base + NSTAGE * 2 * GRAM_TILE_BYTES + warp * (GRAM_STAGING_BYTES / GRAM_CONSUMER_WARPS);
        gram_epilogue_staged<EPI>(acc, p, i0, j0, f1, f2, stg, &tmO, ti * GRAM_BM + 32 * wm, tj * GRAM_BN + 64 * wn, lane, pg, t, Exp2Tab{tab_base + 8u * (lane & 15)});
      }
    } else if (!common) {
      gram_epilogue<EPI, true, 0, MT, NT>(acc, p, i0, j0, rel0, w1, w2, prow, pcol, ldo, Exp2Tab{tab_base + 8u * (lane & 15)});
    } else if (interior && !diag_tile) {
      gram_epilogue<EPI, false, 2, MT, NT>(acc, p, i0, j0, 0, f1, f2m, prow, pcol, ldo, Exp2Tab{tab_base + 8u * (lane & 15)});
    } else {
      gram_epilogue<EPI, true, 2, MT, NT>(acc, p, i0, j0, rel0, w1, w2, prow, pcol, ldo, Exp2Tab{tab_base + 8u * (lane & 15)});
    }
#ifdef PLB_GRAM_TRACE
    if (tr && threadIdx.x == 0) {
      if (tiles_done == 0) tr[3] = gtime();
      ++tiles_done;
      tr[4] = gtime();
      tr[6] = tiles_done;
    }
#endif
  }
  if (TMAST && lane == 0) tma_store_wait_all();  // the staging buffers must outlive the bulk stores that read them
}

}  // namespace plb
