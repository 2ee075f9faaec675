// syrk.cuh — weighted, optionally mean-centred SYRK over rows:  S = sum_r w_r (a_r - m)(a_r - m)'   (Kf x Kf)
// plus, fused on the diagonal tiles, the moment vectors  v = sum_r w_r b_r (a_r - m)  and  c = sum_{r<ne} w_r a_r.
//
// Replaces every normal-equation assembly site of the reference (src/Learning/linear-learn.jl:16-17, 45-49, 94-95,
// 139-143, 157-158, 200, 253, 355-356) and the per-row outer-product mean of compute_eigen
// (src/DimensionReduction/dimension_reduction.jl:12) with the centring of fit!(::PCAState) (pca_state.jl:34-37).
//
// Geometry: the contraction runs over ROWS of one row-major R x Kf matrix, so both MMA operands are "MN-major"
// (the output index is contiguous in memory) — the transpose of the Gram kernel's geometry:
//   * a work unit is 8 warp jobs x all rows. Off-diagonal 128 x 128 tiles are one unit each (4 x 2 warp blocks of 32 x 64).
//     Diagonal tiles are packed two per unit as four 64 x 64 diagonal squares (one warp each, only the 10 upper box pairs of the
//     4 x 4 grid of 16-column fragment boxes) plus their four upper-right 32 x 64 blocks — see syrk_warp_block. A CTA runs a
//     host-built list of PIECES (unit, row range) through one pipeline (SyrkPiece; the schedule is a table, plb200_api.cu).
//     Partials go to a workspace in fragment order (perfectly coalesced) and a deterministic second-stage kernel sums each unit's
//     partials in fixed order, mirrors, and applies scale / ridge / intercept;
//   * stage = 16 rows x 128 columns per operand, brought in as 8 TMA boxes of 16 rows x 16 columns (128 bytes wide,
//     128B swizzle); diagonal tiles load one operand only;
//   * fragment loads are 128-bit: lane (g,t) reads columns (2g, 2g+1) of row 2t+par(+8h) of a box — two m-tiles of
//     one k4-step. The contraction index is permuted (MMA k <-> stage row 2t+par), which together with the 128B
//     swizzle makes every quarter-warp hit 8 distinct 16-byte chunks (conflict-free);
//   * the per-row weight (and the mean subtraction) is applied to the A fragment in registers: the weighted /
//     centred matrix never exists in memory. Centring is ONE-SIDED: S'_ij = sum_r (a_ri - m_i) a_rj is accumulated and the
//     second stage subtracts rho_i m_j with rho_i = sum_r (a_ri - m_i) (accumulated exactly like the moment vectors). This
//     equals sum_r (a_ri - m_i)(a_rj - m_j) without the 8 extra FP64 subtractions per k4-step on the B fragments (scalar
//     FP64 shares the pipe with the DMMAs); the terms are |sigma*m| instead of sigma^2, an error of order eps*(m/sigma)
//     relative to the largest entry in the worst case, far inside the 1e-10 contract even at mean/sigma = 50.
#pragma once
#include <type_traits>

#include "ptx.cuh"

namespace plb {

constexpr int SYRK_BT = 128;      // output tile edge
constexpr int SYRK_BK = 16;       // rows per stage
constexpr int SYRK_STAGES = 5;
constexpr int SYRK_CONSUMER_WARPS = 8;
constexpr int SYRK_THREADS = (SYRK_CONSUMER_WARPS + 4) * 32;  // 2 DMMA warpgroups + 1 producer warpgroup
constexpr int SYRK_BOX_BYTES = SYRK_BK * 16 * 8;        // 2 KB: 16 rows x 16 cols
constexpr int SYRK_OP_BYTES = 8 * SYRK_BOX_BYTES;       // 16 KB per operand per stage
constexpr int SYRK_SMEM_BYTES = SYRK_STAGES * 2 * SYRK_OP_BYTES + 1024 + 256;
constexpr int SYRK_UNIT_ACC = 80;                        // accumulator slots per thread in a unit's partial (squares: 80, blocks: 64)
constexpr int SYRK_TILE_ELEMS = SYRK_UNIT_ACC * 256;     // doubles per (split, unit) partial

// One piece of work: `nstages` pipeline stages (16 rows each) of work unit `unit`, starting at stage `first_stage`; its partial goes to
// workspace slot `slot`. The slots of one unit are contiguous and ordered by row position, so the second-stage sum has a fixed order
// whatever the schedule. A CTA runs a short list of pieces back to back through one TMA/mbarrier pipeline (no refill between pieces).
struct SyrkPiece {
  int unit, first_stage, nstages, slot;
};

struct SyrkParams {
  int64_t R, Kf;
  const double* w;        // per-row weights or NULL
  int64_t ne;             // rows < ne are "energy rows": weight w_e and member of the intercept column
  double w_e, w_f;
  const double* b;        // right-hand side or NULL
  const double* mean;     // column means (centred mode) or NULL
  int T;                  // column blocks = ceil(Kf / 128)
  int ntiles;             // work units: T (T-1) / 2 off-diagonal tiles + ceil(T / 2) diagonal units
  const SyrkPiece* pieces;  // the work list (built on the host, see SyrkPiece): CTA q runs pieces [cta_first[q], cta_first[q + 1])
  const int* cta_first;
  double* ws_tiles;       // [partial slot][80 acc][256 threads]
  int aux_box;            // >= 0: the 16-column box (global index) that is loaded from the auxiliary tensor instead of A (see
                          // syrk_aux_box_kernel: the moment vectors ride along as two extra columns); -1: none
  long long* trace;       // NULL, or 10 int64 per CTA: [globaltimer at start, SM id, end time of consumer warps 0..7] (PLB200_SYRK_TRACE)
};

// Work decomposition of the upper triangle of the T x T grid of 128 x 128 tiles.
//   * off-diagonal tile (ti < tj): one unit, 8 warps x (32 x 64) "blocks" = 32 DMMA tiles per k4-step each;
//   * a diagonal tile needs its upper-right 64 x 64 quarter (two 32 x 64 blocks) and two 64 x 64 diagonal SQUARES. A square is
//     computed by ONE warp at the granularity of the 16-column fragment boxes: the 10 box pairs (bi <= bj) of its 4 x 4 box grid =
//     40 DMMA tiles per k4-step (80 accumulators), instead of the 64 tiles two full blocks would spend — and its A and B
//     fragments are the same registers (4 LDS.128 per 40 DMMAs);
//   * a diagonal UNIT packs two diagonal tiles: warps 0-3 = the four squares, warps 4-7 = the four upper-right blocks, so every
//     SM sub-partition (warp w and w+4) carries 40 + 32 = 72 DMMA tiles per k4-step: 1.125 x an off-diagonal unit for two diagonal
//     tiles. K = 500: 6 + 2 units costing 8.25 (the 6-of-8-blocks packing cost 9, whole diagonal tiles 10); computed / algorithmic
//     flops 1.08 instead of 1.18.
__host__ __device__ __forceinline__ int syrk_num_units(int T) { return T * (T - 1) / 2 + (T + 1) / 2; }

struct SyrkWarpBlock {
  int kind;       // 0: 32 x 64 block, 1: 64 x 64 diagonal square
  int cb_a, wm;   // block: i columns 128 cb_a + 32 wm + [0, 32) ; square: i and j columns 128 cb_a + 64 wm + [0, 64)
  int cb_b, wn;   // block: j columns 128 cb_b + 64 wn + [0, 64)
  int src_a, src_b;  // which of the unit's two operand buffers holds that column block
  bool valid;
};

// work unit u, warp w -> its job. cb0 / cb1: the (at most two) column blocks the unit loads into buffers 0 / 1.
__host__ __device__ __forceinline__ SyrkWarpBlock syrk_warp_block(int T, int u, int w, int& cb0, int& cb1) {
  SyrkWarpBlock B;
  const int noff = T * (T - 1) / 2;
  if (u < noff) {  // off-diagonal tile (ti < tj), row-major over the strict upper triangle
    int ti = 0, rem = u;
    while (rem >= T - 1 - ti) {
      rem -= T - 1 - ti;
      ++ti;
    }
    cb0 = ti;
    cb1 = ti + 1 + rem;
    B.kind = 0;
    B.cb_a = cb0; B.wm = w >> 1; B.src_a = 0;
    B.cb_b = cb1; B.wn = w & 1; B.src_b = 1;
    B.valid = true;
    return B;
  }
  const int q = u - noff;  // diagonal unit q: diagonal tiles 2q and (if it exists) 2q + 1
  const bool two = 2 * q + 1 < T;
  cb0 = 2 * q;
  cb1 = two ? 2 * q + 1 : 2 * q;
  const int ww = w & 3;
  const int sel = ww >> 1;  // which of the unit's tiles
  B.valid = (sel == 0) || two;
  const int t = (sel == 1 && two) ? cb1 : cb0;  // idle warps of an odd last unit shadow tile 2q (their result is dropped)
  B.cb_a = B.cb_b = t;
  B.src_a = B.src_b = (t == cb0) ? 0 : 1;
  if (w < 4) {
    B.kind = 1;
    B.wm = B.wn = ww & 1;  // upper-left / lower-right square
  } else {
    B.kind = 0;
    B.wm = ww & 1;  // rows 0-31 / 32-63 of the tile
    B.wn = 1;       // columns 64-127
  }
  return B;
}

// box pair p = 0..9 of a diagonal square -> (bi, bj), bi <= bj, row-major over the upper triangle of the 4 x 4 box grid
__host__ __device__ __forceinline__ void syrk_square_pair(int p, int& bi, int& bj) {
  bi = p < 4 ? 0 : (p < 7 ? 1 : (p < 9 ? 2 : 3));
  bj = p - (bi == 0 ? 0 : (bi == 1 ? 3 : (bi == 2 ? 5 : 6)));
}

// The consumer loop of one warp. SQUARE = false: 32 x 64 block (A: 2 boxes, B: 4 boxes, 32 DMMA tiles);
// SQUARE = true: 64 x 64 diagonal square (4 boxes serve as A and B, 40 DMMA tiles).
//
// Nothing but LDS + DMMA lives in this loop. Measured with the per-CTA trace hook (scripts/syrk_trace.py): any extra work inside it
// costs far more than its instruction count — scalar FP64 shares the pipe with the DMMAs, and fusing the moment vectors here (as DFMAs
// on the fragments, or as one more DMMA n-tile) made a diagonal unit 1.5 - 1.7 x an off-diagonal one instead of 1.06 x. Hence:
//   * the per-row weight is only multiplied in when the CTA's rows do not all share one weight (NEEDW: a real branch per stage with
//     two instantiations of the body — a select would execute the multiplies always);
//   * the moment vectors are NOT accumulated here: they are weighted column sums, an HBM-rate pass of their own
//     (wcolsum_partial_kernel, prelude.cuh), and the rho of the one-sided centring is the centred column sum (colsum with a shift).
template <bool CENTER, bool SQUARE>
__device__ __forceinline__ void syrk_consume(const SyrkParams& p, const SyrkWarpBlock& WB, const uint32_t smem_base, const uint32_t bar_base,
                                             const uint32_t it0, const int KT, const int64_t r_begin, const int64_t r_end, const int slot, const int lane) {
  constexpr int NA = SQUARE ? 4 : 2;       // fragment boxes of the A operand
  constexpr int NT = SQUARE ? 40 : 32;     // DMMA tiles = accumulator pairs
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (SYRK_STAGES + s); };
  const int g = lane >> 2;
  const int t = lane & 3;
  const int colA0 = WB.cb_a * SYRK_BT + (SQUARE ? 64 : 32) * WB.wm;  // first i column of this warp

  double acc[NT][2];
#pragma unroll
  for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = 0.0;

  // column means of this thread's A fragment columns
  double mi[NA][2];
  if (CENTER) {
#pragma unroll
    for (int bx = 0; bx < NA; ++bx)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int64_t col = (int64_t)colA0 + 16 * bx + 2 * g + e;
        mi[bx][e] = col < p.Kf ? __ldg(p.mean + col) : 0.0;
      }
  }

  // Uniform-weight fast path: when every row of this CTA's range carries the same weight (the two-segment form
  // Q = Diagonal([ws1*1_N; ws2*1_M]) away from the segment boundary, or plain sums), the weight is factored out of the
  // sum and applied once to the accumulators at the end instead of to every A fragment.
  bool uniform = (p.w == nullptr);
  double wscale = 1.0;
  if (uniform) {
    if (r_end <= p.ne) wscale = p.w_e;
    else if (r_begin >= p.ne) wscale = p.w_f;
    else uniform = false;
  }
  // rows of the last stage beyond r_end are zero-filled by TMA (rows_per_split is a multiple of the stage height, so
  // r_end < r_begin + rows_per_split only at the end of the matrix); only the centred mode has to mask them (0 - mean != 0)
  const int64_t r_full_end = r_end - (r_end - r_begin) % SYRK_BK;  // stages starting at or after this row are ragged

  const uint32_t offA = WB.src_a * SYRK_OP_BYTES + (SQUARE ? 4 : 2) * WB.wm * SYRK_BOX_BYTES;
  const uint32_t offB = WB.src_b * SYRK_OP_BYTES + 4 * WB.wn * SYRK_BOX_BYTES;

  // one pipeline stage (16 rows = 4 k4-steps); NEEDW: multiply the per-row weight into the A fragments
  auto stage_body = [&](auto needw_tag, const uint32_t sA, const uint32_t sB, const double (&wv)[2][2]) {
    constexpr bool NEEDW = decltype(needw_tag)::value;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        // byte offset of lane's chunk for step (h, par): row rho = 8h + 2t + par, chunk g ^ (rho & 7) = g ^ (2t + par)
        const uint32_t off = (uint32_t)((8 * h + 2 * t + par) * 128 + ((g ^ (2 * t + par)) << 4));
        double2 a[NA];
#pragma unroll
        for (int bx = 0; bx < NA; ++bx) a[bx] = lds128(sA + bx * SYRK_BOX_BYTES + off);
        const double w = wv[h][par];
        if (SQUARE) {
          // The raw fragments a[] are the B operand; the A operand x = (a - mean) * w is a second copy (weights / centring go on
          // ONE operand only). Without centring and weights the copy is the same registers.
          double2 x[NA];
#pragma unroll
          for (int bx = 0; bx < NA; ++bx) {
            x[bx] = a[bx];
            if (CENTER) {
              x[bx].x -= mi[bx][0];
              x[bx].y -= mi[bx][1];
            }
            if (NEEDW) {
              x[bx].x *= w;
              x[bx].y *= w;
            }
          }
          // box pairs (bi <= bj) of the 4 x 4 box grid, 4 DMMA tiles (e, e') each: tile index (pair*2 + e)*2 + e'
          int pr = 0;
#pragma unroll
          for (int bi = 0; bi < 4; ++bi)
#pragma unroll
            for (int bj = bi; bj < 4; ++bj, ++pr) {
              const double2 xx = x[bi < NA ? bi : 0], yy = a[bj < NA ? bj : 0];
              dmma884(acc[4 * pr + 0][0], acc[4 * pr + 0][1], xx.x, yy.x);
              dmma884(acc[4 * pr + 1][0], acc[4 * pr + 1][1], xx.x, yy.y);
              dmma884(acc[4 * pr + 2][0], acc[4 * pr + 2][1], xx.y, yy.x);
              dmma884(acc[4 * pr + 3][0], acc[4 * pr + 3][1], xx.y, yy.y);
            }
        } else {
          double2 b[4];
#pragma unroll
          for (int bx = 0; bx < 4; ++bx) b[bx] = lds128(sB + bx * SYRK_BOX_BYTES + off);
          if (CENTER) {
#pragma unroll
            for (int bx = 0; bx < NA; ++bx) {
              a[bx].x -= mi[bx][0];
              a[bx].y -= mi[bx][1];
            }
          }
          if (NEEDW) {
#pragma unroll
            for (int bx = 0; bx < NA; ++bx) {
              a[bx].x *= w;
              a[bx].y *= w;
            }
          }
          // m-tile mt = 2 bx + e = columns 16 bx + 2g + e ; n-tile nt = 2 by + e' likewise: tile index mt*8 + nt
#pragma unroll
          for (int bx = 0; bx < 2; ++bx)
#pragma unroll
            for (int by = 0; by < 4; ++by) {
              dmma884(acc[(2 * bx + 0) * 8 + 2 * by + 0][0], acc[(2 * bx + 0) * 8 + 2 * by + 0][1], a[bx].x, b[by].x);
              dmma884(acc[(2 * bx + 0) * 8 + 2 * by + 1][0], acc[(2 * bx + 0) * 8 + 2 * by + 1][1], a[bx].x, b[by].y);
              dmma884(acc[(2 * bx + 1) * 8 + 2 * by + 0][0], acc[(2 * bx + 1) * 8 + 2 * by + 0][1], a[bx].y, b[by].x);
              dmma884(acc[(2 * bx + 1) * 8 + 2 * by + 1][0], acc[(2 * bx + 1) * 8 + 2 * by + 1][1], a[bx].y, b[by].y);
            }
        }
      }
    }
  };

  for (int kt = 0; kt < KT; ++kt) {
    const uint32_t it = it0 + (uint32_t)kt;  // pipeline position: runs on across the pieces of a CTA
    const int s = it % SYRK_STAGES;
    const uint32_t ph = (it / SYRK_STAGES) & 1u;
    const int64_t r0 = r_begin + (int64_t)kt * SYRK_BK;
    const bool need_w = !uniform || (CENTER && r0 >= r_full_end);
    const uint32_t sA = smem_base + s * 2 * SYRK_OP_BYTES + offA;
    const uint32_t sB = smem_base + s * 2 * SYRK_OP_BYTES + offB;
    // per-row weights of the 4 stage rows this lane contracts over: independent of the TMA data, in flight during the wait
    double wv[2][2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        const int64_t r = r0 + 8 * h + 2 * t + par;
        // uniform CTAs apply their one weight to the accumulators at the end (wscale): on the ragged last stage of the centred mode
        // the per-row factor is then only the 1 / 0 mask of the zero-filled rows, not the weight a second time
        double w = 0.0;
        if (need_w && r < r_end) w = uniform ? 1.0 : (p.w ? __ldg(p.w + r) : (r < p.ne ? p.w_e : p.w_f));
        wv[h][par] = w;
      }
    mbar_wait(full_bar(s), ph);
    if (need_w) stage_body(std::true_type{}, sA, sB, wv);
    else stage_body(std::false_type{}, sA, sB, wv);
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar(s));
  }
  if (uniform && wscale != 1.0) {
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      acc[i][0] *= wscale;
      acc[i][1] *= wscale;
    }
  }

  // ---- partial to the workspace in fragment order: [acc index][consumer thread] (coalesced) ----
  double* wst = p.ws_tiles + (int64_t)slot * SYRK_TILE_ELEMS;
  const int ctid = threadIdx.x;  // 0..255
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int c = 0; c < 2; ++c) wst[(i * 2 + c) * 256 + ctid] = acc[i][c];
}

template <bool CENTER>
__global__ void __launch_bounds__(SYRK_THREADS, 1)
syrk_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmx, const SyrkParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int pc0 = p.cta_first[blockIdx.x], pc1 = p.cta_first[blockIdx.x + 1];  // this CTA's pieces
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + SYRK_STAGES * 2 * SYRK_OP_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (SYRK_STAGES + s); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SYRK_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), SYRK_CONSUMER_WARPS);
    }
    fence_barrier_init();
    if (p.trace) {
      long long* tr = p.trace + (int64_t)blockIdx.x * 10;
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      long long t0;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      tr[0] = t0;
      tr[1] = smid;
    }
  }
  __syncthreads();

  if (warp >= SYRK_CONSUMER_WARPS) {
    setmaxnreg_dec<24>();  // 256*240 + 128*24 = 64512 = 384 threads x 168 registers (the CTA pool)
    if (warp == SYRK_CONSUMER_WARPS && lane == 0) {
      tma_prefetch_desc(&tm);
      if (p.aux_box >= 0) tma_prefetch_desc(&tmx);
      uint32_t it = 0;
      for (int pc = pc0; pc < pc1; ++pc) {
        const SyrkPiece pcx = p.pieces[pc];
        int cb0, cb1;
        (void)syrk_warp_block(p.T, pcx.unit, 0, cb0, cb1);
        const bool one_buf = (cb0 == cb1);
        const uint32_t bytes = one_buf ? SYRK_OP_BYTES : 2 * SYRK_OP_BYTES;
        for (int kt = 0; kt < pcx.nstages; ++kt, ++it) {
          const int s = it % SYRK_STAGES;
          const uint32_t ph = (it / SYRK_STAGES) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_arrive_expect_tx(full_bar(s), bytes);
          const uint32_t dst = smem_base + s * 2 * SYRK_OP_BYTES;
          const int r0 = (pcx.first_stage + kt) * SYRK_BK;  // rows beyond R are zero-filled by TMA
#pragma unroll
          for (int bx = 0; bx < 8; ++bx) {
            if (cb0 * 8 + bx == p.aux_box) tma_load_2d(dst + bx * SYRK_BOX_BYTES, &tmx, full_bar(s), 0, r0);
            else tma_load_2d(dst + bx * SYRK_BOX_BYTES, &tm, full_bar(s), cb0 * SYRK_BT + 16 * bx, r0);
          }
          if (!one_buf) {
#pragma unroll
            for (int bx = 0; bx < 8; ++bx) {
              if (cb1 * 8 + bx == p.aux_box) tma_load_2d(dst + SYRK_OP_BYTES + bx * SYRK_BOX_BYTES, &tmx, full_bar(s), 0, r0);
              else tma_load_2d(dst + SYRK_OP_BYTES + bx * SYRK_BOX_BYTES, &tm, full_bar(s), cb1 * SYRK_BT + 16 * bx, r0);
            }
          }
        }
      }
    }
    return;
  }

  // ------------------------------ consumers ------------------------------
  setmaxnreg_inc<240>();
  uint32_t it = 0;
  for (int pc = pc0; pc < pc1; ++pc) {
    const SyrkPiece pcx = p.pieces[pc];
    int cb0, cb1;
    const SyrkWarpBlock WB = syrk_warp_block(p.T, pcx.unit, warp, cb0, cb1);
    const int64_t r_begin = (int64_t)pcx.first_stage * SYRK_BK;
    int64_t r_end = r_begin + (int64_t)pcx.nstages * SYRK_BK;
    if (r_end > p.R) r_end = p.R;
    if (WB.kind == 1) syrk_consume<CENTER, true>(p, WB, smem_base, bar_base, it, pcx.nstages, r_begin, r_end, pcx.slot, lane);
    else syrk_consume<CENTER, false>(p, WB, smem_base, bar_base, it, pcx.nstages, r_begin, r_end, pcx.slot, lane);
    it += (uint32_t)pcx.nstages;
  }
  if (p.trace && lane == 0) {
    long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    p.trace[(int64_t)blockIdx.x * 10 + 2 + warp] = t1;
  }
}

// Second stage: S[i][j] = S[j][i] = sum over the unit's partial slots (fixed order) of the partial blocks; v, c likewise.
// Element (acc index a, consumer thread tid) of work unit u:
//   warp = tid>>5 -> block (cb_a, wm, cb_b, wn) = syrk_warp_block(T, u, warp); g = (tid&31)>>2, t = tid&3
//   mt = a>>4, nt = (a>>1)&7, c = a&1
//   i = 128 cb_a + 32 wm + 16 (mt>>1) + 2 g + (mt&1)
//   j = 128 cb_b + 64 wn + 16 (nt>>1) + 4 t + 2 c + (nt&1)
// Centred mode (mean != NULL): rho (the reduced "c" vector, K doubles, produced by syrk_rho_kernel first) closes the
// one-sided centring: S_ij = S'_ij - rho_i * m_j.
// vec != NULL ("augmented" mode): columns Kf and Kf+1 of the contraction held (b_r) and ([r < ne]) — see syrk_aux_box_kernel — so
//   S_aug[i][Kf] = v_i, S_aug[i][Kf+1] = c_i (i < Kf), S_aug[Kf+1][Kf+1] = s_0, S_aug[Kf][Kf+1] = s_1   ->   vec = [v ; c ; s].
__global__ void syrk_reduce_kernel(const double* __restrict__ ws_tiles, int T, const int* __restrict__ unit_first, int64_t Kf, double* __restrict__ S,
                                   const double* __restrict__ mean, const double* __restrict__ rho, double* __restrict__ vec) {
  const int unit = blockIdx.x;
  const int e = blockIdx.y * blockDim.x + threadIdx.x;  // one partial element per thread: grid = (units, SYRK_TILE_ELEMS / 256)
  if (e >= SYRK_TILE_ELEMS) return;
  const int a = e >> 8;
  const int tid = e & 255;
  const int warp = tid >> 5, g = (tid & 31) >> 2, t = tid & 3;
  int cb0, cb1;
  const SyrkWarpBlock WB = syrk_warp_block(T, unit, warp, cb0, cb1);
  if (!WB.valid) return;
  int64_t i, j;
  if (WB.kind == 0) {
    if (a >= 64) return;
    const int mt = a >> 4, nt = (a >> 1) & 7, c = a & 1;
    i = (int64_t)WB.cb_a * SYRK_BT + 32 * WB.wm + 16 * (mt >> 1) + 2 * g + (mt & 1);
    j = (int64_t)WB.cb_b * SYRK_BT + 64 * WB.wn + 16 * (nt >> 1) + 4 * t + 2 * c + (nt & 1);
  } else {
    int bi, bj;
    syrk_square_pair(a >> 3, bi, bj);
    const int e1 = (a >> 2) & 1, e2 = (a >> 1) & 1, c = a & 1;
    const int64_t base = (int64_t)WB.cb_a * SYRK_BT + 64 * WB.wm;
    i = base + 16 * bi + 2 * g + e1;
    j = base + 16 * bj + 4 * t + 2 * c + e2;
  }
  if (j < i) return;  // diagonal boxes: keep j >= i, mirror (weights are applied to one operand only)
  const int64_t Ka = vec ? Kf + 2 : Kf;
  if (j >= Ka) return;  // (i <= j)
  // four interleaved partial sums (splits sp, sp+1, sp+2, sp+3 mod 4) keep 4 loads in flight; the order is fixed -> deterministic
  const int slot0 = unit_first[unit], splits = unit_first[unit + 1] - slot0;  // this unit's partials: contiguous slots, in row order
  const double* src = ws_tiles + (int64_t)slot0 * SYRK_TILE_ELEMS + e;
  const int64_t stride = SYRK_TILE_ELEMS;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int sp = 0;
  for (; sp + 4 <= splits; sp += 4) {
    s0 += src[(int64_t)sp * stride];
    s1 += src[(int64_t)(sp + 1) * stride];
    s2 += src[(int64_t)(sp + 2) * stride];
    s3 += src[(int64_t)(sp + 3) * stride];
  }
  for (; sp < splits; ++sp) s0 += src[(int64_t)sp * stride];
  double sum = (s0 + s1) + (s2 + s3);
  if (j >= Kf) {  // augmented columns: the moment vectors and the two scalars
    if (i < Kf) vec[(j - Kf) * Kf + i] = sum;
    else if (j == Kf + 1) vec[2 * Kf + (i == Kf ? 1 : 0)] = sum;
    return;
  }
  if (mean) sum = fma(-rho[i], mean[j], sum);
  S[i * Kf + j] = sum;
  S[j * Kf + i] = sum;
}

// The auxiliary 16-column box of the augmented contraction: for every row r
//   aux[r][0 .. rem) = A[r][16 q .. 16 q + rem)   (the last real columns, q = Kf / 16, rem = Kf % 16 <= 14)
//   aux[r][rem] = b_r (0 without a right-hand side),  aux[r][rem + 1] = [r < ne],  the rest 0.
// The SYRK producer loads box q from this tensor instead of A, so A'Wb, the intercept row A'W 1_e and the two scalars fall out of
// the SAME DMMA contraction as two more columns of the (already padded) last column block: no instruction is added to the main loop,
// and no second pass over A is needed (a weighted column-sum pass costs 2.3 ms of HBM time at config 3, 10 % of the SYRK).
// ones != 0 (centred mode): column rem holds 1 instead of b_r, so that S'_aug[i][Kf] = sum_r (a_ri - m_i) = rho_i — exactly the sum of
// the rounded differences the kernel contracted, which is what closes the one-sided centring.
__global__ void syrk_aux_box_kernel(const double* __restrict__ A, int64_t R, int64_t Kf, int64_t lda, const double* __restrict__ b, int64_t ne,
                                    int ones, double* __restrict__ aux) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= R * 16) return;
  const int64_t r = e >> 4;
  const int c = (int)(e & 15);
  const int rem = (int)(Kf & 15);
  double v = 0.0;
  if (c < rem) v = A[r * lda + (Kf - rem) + c];
  else if (c == rem) v = ones ? 1.0 : (b ? b[r] : 0.0);
  else if (c == rem + 1) v = r < ne ? 1.0 : 0.0;
  aux[e] = v;
}

// s[0] = sum_{r<ne} w_r, s[1] = sum_{r<ne} w_r b_r  (one block, fixed-order tree: deterministic)
__global__ void syrk_scalar_kernel(const double* __restrict__ w, const double* __restrict__ b, int64_t ne, double w_e, double* __restrict__ s) {
  __shared__ double sh0[1024], sh1[1024];
  double a0 = 0.0, a1 = 0.0;
  for (int64_t r = threadIdx.x; r < ne; r += blockDim.x) {
    const double wr = w ? w[r] : w_e;
    a0 += wr;
    if (b) a1 = fma(wr, b[r], a1);
  }
  sh0[threadIdx.x] = a0;
  sh1[threadIdx.x] = a1;
  __syncthreads();
  for (int off = blockDim.x >> 1; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) {
      sh0[threadIdx.x] += sh0[threadIdx.x + off];
      sh1[threadIdx.x] += sh1[threadIdx.x + off];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    s[0] = sh0[0];
    s[1] = sh1[0];
  }
}

// (AtWA, AtWb) of order Kf + ic from S, vec = [v ; c ; s]: intercept is row/column 0 (linear-learn.jl:239-241),
// lambda on every diagonal entry (linear-learn.jl:253).
__global__ void normal_eq_finish_kernel(const double* __restrict__ S, const double* __restrict__ vec, int64_t Kf, int ic, double lambda,
                                        double* __restrict__ AtWA, double* __restrict__ AtWb) {
  const int64_t n = Kf + ic;
  const int64_t total = n * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / n, j = e % n;
    double v;
    if (ic && (i == 0 || j == 0)) {
      if (i == 0 && j == 0) v = vec[2 * Kf];
      else v = vec[Kf + (i == 0 ? j - 1 : i - 1)];
    } else {
      // read the j >= i triangle only: after a multi-rank all-reduce S[i][j] and S[j][i] may differ in the last bit
      // (a ring reduces different chunks in different rank orders); the result must be EXACTLY symmetric
      const int64_t a = i - ic, b = j - ic;
      v = a <= b ? S[a * Kf + b] : S[b * Kf + a];
    }
    if (i == j) v += lambda;
    AtWA[e] = v;
  }
  if (AtWb && vec) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      AtWb[i] = (ic && i == 0) ? vec[2 * Kf + 1] : vec[i - ic];
    }
  }
}

// S[i][j] = S[j][i] = S'[i][j] - rho_i m_j over the j >= i triangle: closes the one-sided centring when rho came out of the
// augmented contraction (the reduce kernel has already mirrored S', so both triangles hold S'[min][max])
__global__ void syrk_center_fix_kernel(double* __restrict__ S, int64_t Kf, const double* __restrict__ rho, const double* __restrict__ mean) {
  const int64_t total = Kf * Kf;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / Kf, j = e % Kf;
    if (j < i) continue;
    const double v = fma(-rho[i], mean[j], S[e]);
    S[e] = v;
    S[j * Kf + i] = v;
  }
}

}  // namespace plb
