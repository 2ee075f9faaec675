// syrk.cuh — weighted, optionally mean-centred SYRK over rows:  S = sum_r w_r (a_r - m)(a_r - m)'   (Kf x Kf)
// plus, fused on the diagonal tiles, the moment vectors  v = sum_r w_r b_r (a_r - m)  and  c = sum_{r<ne} w_r a_r.
//
// Replaces every normal-equation assembly site of the reference (src/Learning/linear-learn.jl:16-17, 45-49, 94-95,
// 139-143, 157-158, 200, 253, 355-356) and the per-row outer-product mean of compute_eigen
// (src/DimensionReduction/dimension_reduction.jl:12) with the centring of fit!(::PCAState) (pca_state.jl:34-37).
//
// Geometry: the contraction runs over ROWS of one row-major R x Kf matrix, so both MMA operands are "MN-major"
// (the output index is contiguous in memory) — the transpose of the Gram kernel's geometry:
//   * CTA (unit, split): a work unit is 8 warp blocks of 32 (i) x 64 (j) output elements x a contiguous row range.
//     Off-diagonal 128 x 128 tiles are one unit each (4 x 2 warp blocks). A diagonal tile only needs 6 of its 8 warp
//     blocks (the two below the diagonal are pure mirror images), so the diagonal blocks of ALL tiles are packed 8 per
//     unit (a unit then spans at most two column blocks, the same shared-memory footprint as an off-diagonal tile):
//     T(T-1)/2 + ceil(6T/8) units instead of T(T+1)/2 tiles, e.g. 9 instead of 10 at K = 500, and every warp of every
//     unit does useful work. Partial blocks go to a workspace in fragment order (perfectly coalesced) and a
//     deterministic second-stage kernel sums the splits in fixed order, mirrors, and applies scale / ridge / intercept;
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
constexpr int SYRK_TILE_ELEMS = SYRK_BT * SYRK_BT;      // 16384 accumulators per partial tile

struct SyrkParams {
  int64_t R, Kf;
  const double* w;        // per-row weights or NULL
  int64_t ne;             // rows < ne are "energy rows": weight w_e and member of the intercept column
  double w_e, w_f;
  const double* b;        // right-hand side or NULL
  const double* mean;     // column means (centred mode) or NULL
  int want_vec;           // accumulate v / c on diagonal tiles
  int T;                  // column blocks = ceil(Kf / 128)
  int ntiles;             // work units: T (T-1) / 2 off-diagonal tiles + ceil(6 T / 8) packs of diagonal warp blocks
  int splits;             // row splits
  int64_t rows_per_split; // multiple of SYRK_BK
  double* ws_tiles;       // [split][tile][64 acc][256 threads]
  double* ws_vec;         // [split][2][T*128]   (v then c), only if want_vec
};

// The 6 warp blocks (wm = 32-column slab of i, wn = 64-column slab of j) of a diagonal tile that touch j >= i.
// Blocks 0, 2, 4, 5 — one per wm — also accumulate the moment vectors of their 32 i-columns.
__host__ __device__ __forceinline__ int syrk_num_units(int T) { return T * (T - 1) / 2 + (6 * T + 7) / 8; }

struct SyrkWarpBlock {
  int cb_a, wm;   // i columns: 128 cb_a + 32 wm + [0, 32)
  int cb_b, wn;   // j columns: 128 cb_b + 64 wn + [0, 64)
  int src_a, src_b;  // which of the unit's two operand buffers holds that column block
  bool valid, vec;
};

// work unit u, warp w -> its block. cb0 / cb1: the (at most two) column blocks the unit loads into buffers 0 / 1.
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
    B.cb_a = cb0; B.wm = w >> 1; B.src_a = 0;
    B.cb_b = cb1; B.wn = w & 1; B.src_b = 1;
    B.valid = true;
    B.vec = false;
    return B;
  }
  const int q = u - noff;  // pack q holds diagonal warp blocks 8q .. 8q+7 of the list (tile t, k): index 6t + k
  const int first = 8 * q, last = (8 * q + 7 < 6 * T - 1) ? 8 * q + 7 : 6 * T - 1;
  cb0 = first / 6;
  cb1 = last / 6;
  const int idx = first + w;
  B.valid = idx < 6 * T;
  const int t = B.valid ? idx / 6 : cb0, k = B.valid ? idx % 6 : 0;
  // k: 0 -> (0,0)  1 -> (0,1)  2 -> (1,0)  3 -> (1,1)  4 -> (2,1)  5 -> (3,1)
  B.wm = k < 4 ? (k >> 1) : (k - 2);
  B.wn = k < 4 ? (k & 1) : 1;
  B.cb_a = B.cb_b = t;
  B.src_a = B.src_b = (t == cb0) ? 0 : 1;
  B.vec = B.valid && (k == 0 || k == 2 || k == 4 || k == 5);
  return B;
}

template <bool CENTER, bool VEC>
__global__ void __launch_bounds__(SYRK_THREADS, 1)
syrk_kernel(const __grid_constant__ CUtensorMap tm, const SyrkParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + SYRK_STAGES * 2 * SYRK_OP_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (SYRK_STAGES + s); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x;  // work unit
  const int split = blockIdx.y;
  int cb0, cb1;
  const SyrkWarpBlock WB = syrk_warp_block(p.T, tile, warp < SYRK_CONSUMER_WARPS ? warp : 0, cb0, cb1);
  const bool one_buf = (cb0 == cb1);
  const int64_t r_begin = (int64_t)split * p.rows_per_split;
  int64_t r_end = r_begin + p.rows_per_split;
  if (r_end > p.R) r_end = p.R;
  const int KT = r_end > r_begin ? (int)((r_end - r_begin + SYRK_BK - 1) / SYRK_BK) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SYRK_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), SYRK_CONSUMER_WARPS);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp >= SYRK_CONSUMER_WARPS) {
    setmaxnreg_dec<24>();  // 256*240 + 128*24 = 64512 = 384 threads x 168 registers (the CTA pool)
    if (warp == SYRK_CONSUMER_WARPS && lane == 0) {
      tma_prefetch_desc(&tm);
      const uint32_t bytes = one_buf ? SYRK_OP_BYTES : 2 * SYRK_OP_BYTES;
      for (int kt = 0; kt < KT; ++kt) {
        const int s = kt % SYRK_STAGES;
        const uint32_t ph = (kt / SYRK_STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_arrive_expect_tx(full_bar(s), bytes);
        const uint32_t dst = smem_base + s * 2 * SYRK_OP_BYTES;
        const int r0 = (int)(r_begin + (int64_t)kt * SYRK_BK);
        // NOTE: the last stage of a split may read rows of the next split (r >= r_end); they are zero-weighted below.
#pragma unroll
        for (int bx = 0; bx < 8; ++bx) tma_load_2d(dst + bx * SYRK_BOX_BYTES, &tm, full_bar(s), cb0 * SYRK_BT + 16 * bx, r0);
        if (!one_buf) {
#pragma unroll
          for (int bx = 0; bx < 8; ++bx)
            tma_load_2d(dst + SYRK_OP_BYTES + bx * SYRK_BOX_BYTES, &tm, full_bar(s), cb1 * SYRK_BT + 16 * bx, r0);
        }
      }
    }
    return;
  }

  // ------------------------------ consumers ------------------------------
  setmaxnreg_inc<240>();
  const int wm = WB.wm;  // 32-column slab of the i (row of S) index inside column block WB.cb_a
  const int wn = WB.wn;  // 64-column slab of the j index inside column block WB.cb_b
  const int g = lane >> 2;
  const int t = lane & 3;

  double acc[4][8][2];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

  // moment vectors (one diagonal warp block per 32-column slab): columns i = 128 cb_a + 32 wm + 16 bx + 2g + e
  double vb[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  double vc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  const bool do_vec = VEC && WB.vec;

  // column means of this thread's fragment columns
  double mi[2][2];
  if (CENTER) {
#pragma unroll
    for (int bx = 0; bx < 2; ++bx)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int64_t col = (int64_t)WB.cb_a * SYRK_BT + 32 * wm + 16 * bx + 2 * g + e;
        mi[bx][e] = col < p.Kf ? __ldg(p.mean + col) : 0.0;
      }
  }

  // Uniform-weight fast path: when every row of this CTA's range carries the same weight (the two-segment form
  // Q = Diagonal([ws1*1_N; ws2*1_M]) away from the segment boundary, or plain sums), the weight is factored out of the
  // sum and applied once to the 64 accumulators at the end instead of to every A fragment: the FP64 multiplies would
  // otherwise queue behind the DMMAs on the shared FP64 pipe inside the main loop.
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

  // byte offset of lane's chunk for step (h, par): row rho = 8h + 2t + par, chunk g ^ (rho & 7)
  // rho & 7 = 2t + par  ->  (g ^ (2t+par)) << 4 ;  row offset rho * 128
  for (int kt = 0; kt < KT; ++kt) {
    const int s = kt % SYRK_STAGES;
    const uint32_t ph = (kt / SYRK_STAGES) & 1u;
    const int64_t r0 = r_begin + (int64_t)kt * SYRK_BK;
    const bool need_w = !uniform || (CENTER && r0 >= r_full_end);

    // per-row weights / rhs of the 4 stage rows this lane contracts over (independent of the TMA data)
    double wv[2][2], bv[2][2], cv[2][2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        const int64_t r = r0 + 8 * h + 2 * t + par;
        const bool ok = r < r_end;
        double w = 0.0;
        if (ok) w = uniform ? 1.0 : (p.w ? __ldg(p.w + r) : (r < p.ne ? p.w_e : p.w_f));
        wv[h][par] = w;
        if (VEC) {
          bv[h][par] = (ok && p.b) ? __ldg(p.b + r) : 0.0;
          cv[h][par] = (ok && (CENTER || r < p.ne)) ? 1.0 : 0.0;  // centred mode: rho = sum over ALL rows of (a - m)
        }
      }

    mbar_wait(full_bar(s), ph);
    const uint32_t sA = smem_base + s * 2 * SYRK_OP_BYTES + WB.src_a * SYRK_OP_BYTES + (2 * wm) * SYRK_BOX_BYTES;
    const uint32_t sB = smem_base + s * 2 * SYRK_OP_BYTES + WB.src_b * SYRK_OP_BYTES + (4 * wn) * SYRK_BOX_BYTES;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        const uint32_t off = (uint32_t)((8 * h + 2 * t + par) * 128 + ((g ^ (2 * t + par)) << 4));
        double2 a[2], b[4];
#pragma unroll
        for (int bx = 0; bx < 2; ++bx) a[bx] = lds128(sA + bx * SYRK_BOX_BYTES + off);
#pragma unroll
        for (int bx = 0; bx < 4; ++bx) b[bx] = lds128(sB + bx * SYRK_BOX_BYTES + off);
        if (CENTER) {
#pragma unroll
          for (int bx = 0; bx < 2; ++bx) {
            a[bx].x -= mi[bx][0];
            a[bx].y -= mi[bx][1];
          }
        }
        if (need_w) {
          const double w = wv[h][par];
#pragma unroll
          for (int bx = 0; bx < 2; ++bx) {
            a[bx].x *= w;
            a[bx].y *= w;
          }
        }
        if (VEC) {
          if (do_vec) {
            const double bb = bv[h][par], cc = cv[h][par];
#pragma unroll
            for (int bx = 0; bx < 2; ++bx) {
              if (!CENTER) {  // centred mode has no right-hand side: only rho (the c slot) is accumulated
                vb[bx][0] = fma(a[bx].x, bb, vb[bx][0]);
                vb[bx][1] = fma(a[bx].y, bb, vb[bx][1]);
              }
              vc[bx][0] = fma(a[bx].x, cc, vc[bx][0]);
              vc[bx][1] = fma(a[bx].y, cc, vc[bx][1]);
            }
          }
        }
        // m-tile (bx, e) = columns 16 bx + 2g + e ; n-tile (bx', e') likewise: 4 x 8 DMMA tiles per k4-step
#pragma unroll
        for (int bx = 0; bx < 2; ++bx)
#pragma unroll
          for (int by = 0; by < 4; ++by) {
            dmma884(acc[2 * bx + 0][2 * by + 0][0], acc[2 * bx + 0][2 * by + 0][1], a[bx].x, b[by].x);
            dmma884(acc[2 * bx + 0][2 * by + 1][0], acc[2 * bx + 0][2 * by + 1][1], a[bx].x, b[by].y);
            dmma884(acc[2 * bx + 1][2 * by + 0][0], acc[2 * bx + 1][2 * by + 0][1], a[bx].y, b[by].x);
            dmma884(acc[2 * bx + 1][2 * by + 1][0], acc[2 * bx + 1][2 * by + 1][1], a[bx].y, b[by].y);
          }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar(s));
  }
  if (uniform && wscale != 1.0) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        acc[mt][nt][0] *= wscale;
        acc[mt][nt][1] *= wscale;
      }
#pragma unroll
    for (int bx = 0; bx < 2; ++bx)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        vb[bx][e] *= wscale;
        vc[bx][e] *= wscale;
      }
  }

  // ---- partial tile to the workspace in fragment order: [acc index][consumer thread] (coalesced) ----
  double* wst = p.ws_tiles + ((int64_t)split * p.ntiles + tile) * SYRK_TILE_ELEMS;
  const int ctid = threadIdx.x;  // 0..255
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int c = 0; c < 2; ++c) wst[((mt * 8 + nt) * 2 + c) * 256 + ctid] = acc[mt][nt][c];

  if (VEC) {
    if (do_vec) {
      // reduce over the 4 lanes t = 0..3 that share g (they hold disjoint stage rows)
#pragma unroll
      for (int bx = 0; bx < 2; ++bx)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double x = vb[bx][e], y = vc[bx][e];
          x += __shfl_xor_sync(0xffffffffu, x, 1);
          x += __shfl_xor_sync(0xffffffffu, x, 2);
          y += __shfl_xor_sync(0xffffffffu, y, 1);
          y += __shfl_xor_sync(0xffffffffu, y, 2);
          if (t == 0) {
            const int col = WB.cb_a * SYRK_BT + 32 * wm + 16 * bx + 2 * g + e;
            double* wv2 = p.ws_vec + (int64_t)split * 2 * p.T * SYRK_BT;
            wv2[col] = x;
            wv2[p.T * SYRK_BT + col] = y;
          }
        }
    }
  }
}

// Second stage: S[i][j] = S[j][i] = sum over splits (fixed order) of the partial blocks; v, c likewise.
// Element (acc index a, consumer thread tid) of work unit u:
//   warp = tid>>5 -> block (cb_a, wm, cb_b, wn) = syrk_warp_block(T, u, warp); g = (tid&31)>>2, t = tid&3
//   mt = a>>4, nt = (a>>1)&7, c = a&1
//   i = 128 cb_a + 32 wm + 16 (mt>>1) + 2 g + (mt&1)
//   j = 128 cb_b + 64 wn + 16 (nt>>1) + 4 t + 2 c + (nt&1)
// Centred mode (mean != NULL): rho (the reduced "c" vector, K doubles, produced by syrk_rho_kernel first) closes the
// one-sided centring: S_ij = S'_ij - rho_i * m_j.
__global__ void syrk_reduce_kernel(const double* __restrict__ ws_tiles, const double* __restrict__ ws_vec, int T, int ntiles,
                                   int splits, int64_t Kf, double* __restrict__ S, double* __restrict__ vec, int want_vec,
                                   const double* __restrict__ mean, const double* __restrict__ rho) {
  const int unit = blockIdx.x;
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < SYRK_TILE_ELEMS; e += gridDim.y * blockDim.x) {
    const int a = e >> 8;
    const int tid = e & 255;
    const int warp = tid >> 5, g = (tid & 31) >> 2, t = tid & 3;
    int cb0, cb1;
    const SyrkWarpBlock WB = syrk_warp_block(T, unit, warp, cb0, cb1);
    if (!WB.valid) continue;
    const int mt = a >> 4, nt = (a >> 1) & 7, c = a & 1;
    const int64_t i = (int64_t)WB.cb_a * SYRK_BT + 32 * WB.wm + 16 * (mt >> 1) + 2 * g + (mt & 1);
    const int64_t j = (int64_t)WB.cb_b * SYRK_BT + 64 * WB.wn + 16 * (nt >> 1) + 4 * t + 2 * c + (nt & 1);
    if (i >= Kf || j >= Kf) continue;
    if (j < i) continue;  // diagonal blocks: keep j >= i, mirror (weights are applied to one operand only)
    double sum = 0.0;
    for (int sp = 0; sp < splits; ++sp) sum += ws_tiles[((int64_t)sp * ntiles + unit) * SYRK_TILE_ELEMS + e];
    if (mean) sum = fma(-rho[i], mean[j], sum);
    S[i * Kf + j] = sum;
    S[j * Kf + i] = sum;
  }
  if (want_vec && blockIdx.y == 0) {
    // every column block's moment vectors: handled by the first T units' blocks (any mapping works: one writer per column)
    for (int64_t col = (int64_t)unit * blockDim.x + threadIdx.x; col < Kf; col += (int64_t)gridDim.x * blockDim.x) {
      double sv = 0.0, sc = 0.0;
      for (int sp = 0; sp < splits; ++sp) {
        const double* wv2 = ws_vec + (int64_t)sp * 2 * T * SYRK_BT;
        sv += wv2[col];
        sc += wv2[(int64_t)T * SYRK_BT + col];
      }
      vec[col] = sv;
      vec[Kf + col] = sc;
    }
  }
}

// rho[col] = sum over splits (fixed order) of the "c" moment vector: rho_i = sum_r (a_ri - m_i) in centred mode
__global__ void syrk_rho_kernel(const double* __restrict__ ws_vec, int T, int splits, int64_t Kf, double* __restrict__ rho) {
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= Kf) return;
  double s = 0.0;
  for (int sp = 0; sp < splits; ++sp) s += ws_vec[(int64_t)sp * 2 * T * SYRK_BT + (int64_t)T * SYRK_BT + col];
  rho[col] = s;
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

}  // namespace plb
