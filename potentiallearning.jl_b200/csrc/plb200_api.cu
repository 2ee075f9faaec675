// plb200_api.cu — the C ABI of libplb200.so (include/plb200.h): context, workspaces, TMA descriptors, launches.
// Everything here is plumbing around the three kernel families in gram.cuh / syrk.cuh / prelude.cuh.
// There is deliberately no CPU code path: every entry point needs an sm_100 device.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/plb200.h"
#include "gram.cuh"
#include "hoststage.h"
#include "prelude.cuh"
#include "syrk.cuh"
#include "eigh.cuh"

using namespace plb;

namespace {

thread_local char g_err[1024] = "";
bool staging_for(const void* host_ptr, size_t bytes) {
  return bytes >= ((size_t)4 << 20) && !getenv("PLB200_NO_STAGING") && HostStager::is_pageable(host_ptr);
}
std::mutex g_mu;
std::atomic<int64_t> g_launches{0};

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      cudaGetLastError();                                                                               \
      return set_err(_e == cudaErrorMemoryAllocation ? PL_ENOMEM : PL_ECUDA, "%s failed: %s (%s:%d)", #expr, \
                     cudaGetErrorString(_e), __FILE__, __LINE__);                                       \
    }                                                                                                   \
  } while (0)
#define PL_TRY(expr)         \
  do {                       \
    int _r = (expr);         \
    if (_r != PL_OK) return _r; \
  } while (0)
#define LAUNCH_CHECK()                  \
  do {                                  \
    g_launches.fetch_add(1);            \
    CUDA_TRY(cudaGetLastError());       \
  } while (0)

enum WsSlot {
  WS_COLSUM_PARTIAL = 0,
  WS_MEAN,
  WS_XC1,
  WS_XC2,
  WS_Y1,
  WS_Y2,
  WS_STAT1,
  WS_STAT2,
  WS_PAD1,
  WS_PAD2,
  WS_PADC,
  WS_SYRK_TILES,
  WS_SYRK_VEC,
  WS_SYRK_RHO,
  WS_H_IN1,
  WS_H_IN2,
  WS_H_IN3,
  WS_H_IN4,
  WS_H_OUT1,
  WS_H_OUT2,
  WS_H_S,
  WS_H_VEC,
  WS_FEAT_L,
  WS_FEAT_OFF,
  WS_FEAT_X,
  WS_NE_BUF0,
  WS_NE_BUF1,
  WS_NE_SACC,
  WS_NE_VACC,
  WS_EIG_A,
  WS_EIG_W,
  WS_EIG_WORK,
  WS_EIG_INFO,
  WS_EIG_GAMMA,
  WS_EIG_P,
  WS_EIG_IDX,
  WS_EIG_SEL,
  WS_TRACE,
  WS_SYRK_SCHED,
  WS_SYRK_AUX,
  WS_ACC_S,
  WS_ACC_V,
  WS_KSD_G1,
  WS_KSD_G2,
  WS_KSD_G3,
  WS_KSD_H,
  WS_KSD_S,
  WS_KSD_PART,
  WS_COV_Z,
  WS_COV_RHO,
  WS_COV_DELTA,
  WS_DPP_P,
  WS_DPP_F,
  WS_DPP_VEC,
  WS_DPP_TMP,
  WS_DPP_ITEMS,
  WS_DPP_U,
  WS_DPP_K,
  WS_DPP_C,
  WS_ESP_T,
  WS_ESP_AUX,
  WS_ELL_K,
  WS_POS,
  WS_POS_C,
  WS_POS_D,
  WS_MET,
  WS_EIGH_S,
  WS_EIGH_QA,
  WS_EIGH_QB,
  WS_EIGH_VT,
  WS_EIGH_SMALL,
  WS_EIGH_TAB,
  WS_NSLOTS
};

// the SYRK schedule of the last shape stays on the device (same shape again: no rebuild, no upload); see build_syrk_schedule
struct SyrkSchedule {
  std::vector<int> cta_first, unit_first;
  std::vector<SyrkPiece> pieces;
  int nctas = 0, nslots = 0;
};
struct SyrkSchedCache {
  int64_t R = -1;
  int T = -1;
  const void* dev = nullptr;
  cudaStream_t up_stream = nullptr;  // the stream the table was uploaded on; other streams wait for `ready` before using it
  cudaEvent_t ready = nullptr;
  SyrkSchedule sch;
};

// One context per GPU the library drives. pl_init binds context 0 to one device (one process per GPU); pl_init_all creates one per
// device of the box (single process, what a Julia session uses). Every internal routine works on the calling thread's CURRENT
// context `g`: the entry points run on context 0, the multi-GPU entry points hand each device's share to a host thread whose
// current context is that device's.
constexpr int PL_MAX_DEV = 16;
struct Ctx {
  bool ready = false;
  int device = -1;
  int index = 0;  // position in g_ctx
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // D2H of finished row blocks overlaps the remaining Gram launches
  cudaEvent_t chunk_ev[16] = {};
  cudaEvent_t ev[6] = {};
  PFN_cuTensorMapEncodeTiled encode = nullptr;
  void* ws[WS_NSLOTS] = {};
  size_t ws_bytes[WS_NSLOTS] = {};
  double t_kernel = 0, t_h2d = 0, t_d2h = 0, t_coll = 0, t_host = 0;
  std::chrono::steady_clock::time_point t_entry;  // when the current host entry point started (need_ready)
  bool profiling = false;
  bool main_timed = false;
  uint64_t ws_gen = 0;  // bumped whenever a workspace slot is (re)allocated or freed: cached CUDA graphs hold workspace pointers
  HostStager stager;  // pinned staging ring + host threads for pageable host memory (hoststage.h)
  SyrkSchedCache sched;
  void* exp_tab = nullptr;  // 2^(j/256) table of the RBF epilogue (gram.cuh exp2_tab), written once at context creation
  void* comm = nullptr;  // ncclComm_t of this device (pl_init_all: one per context; pl_comm_init_rank: context 0)
};
Ctx g_ctx[PL_MAX_DEV];
int g_ndev = 0;  // contexts in use (0 before pl_init)
thread_local Ctx* tl_ctx = &g_ctx[0];
#define g (*tl_ctx)
#define g_stager (g.stager)
#define g_sched (g.sched)

// ev[4], ev[5] bracket the dominant (DMMA) kernel of the last operation when profiling is on
#define MAIN_KERNEL_BEGIN(st) do { if (g.profiling) CUDA_TRY(cudaEventRecord(g.ev[4], st)); } while (0)
#define MAIN_KERNEL_END(st) do { if (g.profiling) { CUDA_TRY(cudaEventRecord(g.ev[5], st)); g.main_timed = true; } } while (0)

int ws_get(int slot, size_t bytes, void** out) {
  if (bytes == 0) bytes = 16;
  if (g.ws_bytes[slot] < bytes) {
    if (g.ws[slot]) {
      CUDA_TRY(cudaFree(g.ws[slot]));
      g.ws[slot] = nullptr;
      g.ws_bytes[slot] = 0;
    }
    // grow with 12.5% slack, 256-byte granularity
    size_t want = (bytes + bytes / 8 + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&g.ws[slot], want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      want = (bytes + 255) & ~(size_t)255;
      e = cudaMalloc(&g.ws[slot], want);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      g.ws[slot] = nullptr;
      return set_err(PL_ENOMEM, "device allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
    }
    g.ws_bytes[slot] = want;
    ++g.ws_gen;
  }
  *out = g.ws[slot];
  return PL_OK;
}

// (multi.inl) the bucket an operation accumulates into: the peer window's when the exchange runs over peer memory, else workspace `slot`
int bucket_get(size_t count, int slot, double** out);

int need_ready() {
  if (!g.ready) return set_err(PL_ENODEV, "plb200 is not initialised: call pl_init (needs an sm_100 GPU; there is no CPU fallback)");
  // every entry point (host and _dev) works on the device of the calling thread's context, whatever device the caller had current
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != g.device) CUDA_TRY(cudaSetDevice(g.device));
  g.t_entry = std::chrono::steady_clock::now();
  return PL_OK;
}

int make_tmap_2d(CUtensorMap* tm, const double* ptr, int64_t dim0, int64_t dim1, int64_t ld, uint32_t box0, uint32_t box1) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) || (ld & 1)) return set_err(PL_EINVAL, "TMA operand must be 16-byte aligned with an even leading dimension");
  cuuint64_t gdim[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g.encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(PL_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (dims %lld x %lld, ld %lld)", (int)r, (long long)dim0, (long long)dim1, (long long)ld);
  return PL_OK;
}

__global__ void pad_copy_kernel(const double* __restrict__ src, int64_t rows, int64_t cols, int64_t lds, double* __restrict__ dst, int64_t ldd) {
  const int64_t total = rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / cols, c = e % cols;
    dst[r * ldd + c] = src[r * lds + c];
  }
}

// returns an operand that TMA / 128-bit loads can read: the original if aligned with even ld, else a padded copy
int tma_operand(const double* src, int64_t rows, int64_t cols, int64_t ld, int slot, cudaStream_t st, const double** out, int64_t* ldout) {
  if (!(reinterpret_cast<uintptr_t>(src) & 15u) && !(ld & 1)) {
    *out = src;
    *ldout = ld;
    return PL_OK;
  }
  const int64_t ldp = (cols + 1) & ~(int64_t)1;
  void* buf = nullptr;
  PL_TRY(ws_get(slot, (size_t)rows * ldp * 8, &buf));
  const int64_t total = rows * cols;
  int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  if (blocks < 1) blocks = 1;
  pad_copy_kernel<<<blocks, 256, 0, st>>>(src, rows, cols, ld, (double*)buf, ldp);
  LAUNCH_CHECK();
  *out = (const double*)buf;
  *ldout = ldp;
  return PL_OK;
}

int64_t colsum_rows_per_block(int64_t R, int64_t Kf) {
  // rows per block: aim at >= 16 blocks per SM, between 32 and 512 rows each (fixed for a given shape: deterministic)
  const int64_t colblocks = ((Kf + 1) / 2 + 127) / 128;
  int64_t rpb = R * colblocks / ((int64_t)g.num_sms * 16);
  rpb = rpb < 32 ? 32 : (rpb > COLSUM_MAX_ROWS_PER_BLOCK ? COLSUM_MAX_ROWS_PER_BLOCK : (rpb + 3) / 4 * 4);
  const int64_t min_rpb = ((R + 65534) / 65535 + 3) / 4 * 4;  // gridDim.y is limited to 65535 row blocks
  return rpb < min_rpb ? min_rpb : rpb;
}

int colsum(const double* dA, int64_t R, int64_t Kf, int64_t lda, double denom, double* dout, cudaStream_t st, const double* dshift = nullptr) {
  const int64_t rpb = colsum_rows_per_block(R, Kf);
  const int64_t nblk = (R + rpb - 1) / rpb;
  void* part = nullptr;
  PL_TRY(ws_get(WS_COLSUM_PARTIAL, (size_t)(nblk > 0 ? nblk : 1) * Kf * 8, &part));
  if (nblk > 0) {
    const int64_t pairs = (Kf + 1) / 2;
    dim3 grid((unsigned)((pairs + 127) / 128), (unsigned)nblk);
    colsum_partial_kernel<<<grid, 128, 0, st>>>(dA, R, Kf, lda, (int)rpb, (double*)part, dshift);
    LAUNCH_CHECK();
  }
  colsum_final_kernel<<<(unsigned)((Kf + 31) / 32), 256, 0, st>>>((const double*)part, nblk, Kf, denom, dout);
  LAUNCH_CHECK();
  return PL_OK;
}

// moment vectors of the normal equations: dv (may be NULL iff db is NULL) = A'(w.b), dc = A'(w.[r < ne]); unscaled, deterministic
int moment_vectors(const double* dA, int64_t R, int64_t Kf, int64_t lda, const double* dw, int64_t ne, double w_e, double w_f, const double* db,
                   double* dv, double* dc, cudaStream_t st) {
  // without a right-hand side only the first ne rows contribute
  const int64_t Rv = db ? R : ne;
  const int64_t rpb = colsum_rows_per_block(Rv, Kf);
  const int64_t nblk = (Rv + rpb - 1) / rpb;
  void* part = nullptr;
  PL_TRY(ws_get(WS_COLSUM_PARTIAL, (size_t)(nblk > 0 ? nblk : 1) * Kf * 8 * 2, &part));
  double* pv = (double*)part;
  double* pc = pv + (nblk > 0 ? nblk : 1) * Kf;
  if (nblk > 0) {
    const int64_t pairs = (Kf + 1) / 2;
    dim3 grid((unsigned)((pairs + 127) / 128), (unsigned)nblk);
    wcolsum_partial_kernel<<<grid, 128, 0, st>>>(dA, Rv, Kf, lda, dw, ne, w_e, w_f, db, (int)rpb, pv, pc);
    LAUNCH_CHECK();
  }
  // blocks beyond the energy rows hold zeros in pc: only the first ceil(ne / rpb) blocks are summed
  const int64_t nblk_c = ne > 0 ? (ne + rpb - 1) / rpb : 0;
  if (db) {
    colsum_final_kernel<<<(unsigned)((Kf + 31) / 32), 256, 0, st>>>(pv, nblk, Kf, 0.0, dv);
    LAUNCH_CHECK();
  } else if (dv) {
    CUDA_TRY(cudaMemsetAsync(dv, 0, (size_t)Kf * 8, st));
  }
  colsum_final_kernel<<<(unsigned)((Kf + 31) / 32), 256, 0, st>>>(pc, nblk_c < nblk ? nblk_c : nblk, Kf, 0.0, dc);
  LAUNCH_CHECK();
  return PL_OK;
}

struct GramLaunch {
  int epi;
  const double *A, *B;
  int64_t nA, ldA, nB, ldB, d;
  const double *ra, *rb;
  double* out;
  int64_t ldo;
  int symmetric, same_operand, fill, packed, part, nparts, alpha;
  double s, beta;
  int tcol_begin = -1, tcol_end = -1;  // column-range enumeration (host streaming path)
  int pdl = 0;                         // the launch directly follows the prelude kernel it depends on: programmatic dependent launch
  int b_shift = 0;                     // RBF (augmented operands, prelude.cuh rbf_aug_kernel): the B operand's last k-step sits b_shift columns further right
};

// <<<>>> with the programmatic-dependent-launch attribute when pdl is set (the kernel then calls pdl_wait() before it touches what its
// predecessor in the stream produces)
template <class... KArgs, class... Args>
cudaError_t launch_ex(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Tile shape of a dense (not tile-packed, not column-range) Gram launch: 64 x 64 tiles when the 128 x 128 tiles of the result would
// leave SMs idle or give too few waves to hide the tail (4 x as many CTAs): N = 1000 (config 1, the kDPP example shape) has 36
// upper-triangle tiles of 128 for 148 SMs, 136 of 64. Measured (profiles/r02_small_gram_probe.json, CUDA-graph replay; r02_skinny_tile_probe.json):
// N = 1000, d = 26 DotProduct 15.1 -> 6.0 us, RBF 30.2 -> 13.9 us; N = 4000, d = 26 38 -> 33 us / 72 -> 52 us; N = 8000 RBF 0.187 -> 0.161 ms
// but DotProduct 0.103 -> 0.111 ms (the TMA-store epilogue of the 128 tile wins once the machine is full); N >= 20 000: 128 x 128 always.
// Hence: 64 x 64 below 4 waves of 128-tiles, below 16 waves for the exp epilogue.
// The host paths that copy a triangle back as a staircase of row blocks use the same edge (the unused half of a DIAGONAL tile is
// written as zeros by the kernel, off-diagonal tiles of the other triangle are not touched).
bool gram_small_tiles(int64_t nA, int64_t nB, bool symmetric, int nparts, int kernel_id) {
  const int T1 = (int)((nA + GRAM_BM - 1) / GRAM_BM), T2 = (int)((nB + GRAM_BN - 1) / GRAM_BN);
  const int64_t t128 = symmetric ? (int64_t)T1 * (T1 + 1) / 2 : (int64_t)T1 * T2;
  const int64_t waves = kernel_id == PL_KERNEL_RBF ? 16 : 4;
  bool small = (t128 + nparts - 1) / nparts < waves * g.num_sms;
  if (const char* e = getenv("PLB200_GRAM_SMALL")) small = e[0] == '1' ? true : (e[0] == '0' ? false : small);
  return small;
}

int launch_gram(const GramLaunch& L, cudaStream_t st) {
  if (L.nA == 0 || L.nB == 0) return PL_OK;
  // rectangular results with few columns (projections, cross kernels against a handful of points): the 128 x 32 tile
  const bool narrow = !L.symmetric && !L.packed && L.nB <= 96;
  const bool small = !narrow && !L.packed && L.tcol_begin < 0 && gram_small_tiles(L.nA, L.nB, L.symmetric, L.nparts, L.epi == EPI_RBF ? PL_KERNEL_RBF : PL_KERNEL_DOT);
  const int BM = small ? 64 : GRAM_BM;
  const int BN = narrow ? 32 : (small ? 64 : GRAM_BN);
  CUtensorMap tmA, tmB;
  PL_TRY(make_tmap_2d(&tmA, L.A, L.d + L.b_shift, L.nA, L.ldA, GRAM_BK, BM));
  PL_TRY(make_tmap_2d(&tmB, L.B, L.d + L.b_shift, L.nB, L.ldB, GRAM_BK, BN));
  GramParams p;
  p.b_shift = L.b_shift;
  p.exp_tab = (const double*)g.exp_tab;
#ifdef PLB_GRAM_TRACE
  p.trace = nullptr;
  if (getenv("PLB200_GRAM_TRACE")) {  // developer build: per-CTA timeline of the MAIN launch appended to that file (scripts/gram_trace.py)
    void* trb;
    PL_TRY(ws_get(WS_TRACE, (size_t)g.num_sms * 8 * 8, &trb));
    CUDA_TRY(cudaMemsetAsync(trb, 0, (size_t)g.num_sms * 8 * 8, st));
    p.trace = (long long*)trb;
  }
#endif

  const int tab_smem = L.epi == EPI_RBF ? EXP_TAB_BYTES : 0;
  p.n1 = L.nA;
  p.n2 = L.nB;
  p.d = L.d;
  p.out = L.out;
  p.ldo = L.ldo;
  p.ra = L.ra;
  p.rb = L.rb;
  p.symmetric = L.symmetric;
  p.same_operand = L.same_operand;
  p.fill = L.fill;
  if (getenv("PLB200_GRAM_NOSTORE") && L.symmetric && !L.packed) p.fill = 0;  // developer experiment: epilogue without off-diagonal stores
  p.packed = L.packed;
  p.alpha = L.alpha;
  p.s = L.s;
  p.beta = L.beta;
  p.bmul = (L.epi == EPI_RBF && L.beta > 0.0) ? 1.0 : L.beta;
  p.T1 = (int)((L.nA + BM - 1) / BM);
  p.T2 = (int)((L.nB + BN - 1) / BN);
  p.tcol_begin = L.tcol_begin;
  p.tcol_end = L.tcol_end;
  p.ntiles_total = L.tcol_begin >= 0 ? gram_colrange_num_slots(L.tcol_begin, L.tcol_end) : gram_num_slots(p.T1, p.T2, p.symmetric);
  p.part = L.part;
  p.nparts = L.nparts;
  const bool pdl = L.pdl && !getenv("PLB200_NO_PDL");
  p.pdl = pdl ? 1 : 0;
  p.tail_first = -1;
  p.tail_count = 0;
  p.T1p = p.T1;
  p.T2p = p.T2;
  const int64_t mine = (p.ntiles_total - L.part + L.nparts - 1) / L.nparts;
  if (mine <= 0) return PL_OK;
  // Tile-packed shards (one GPU's share of a matrix): when the share is only a few waves of CTAs, the last PARTIAL wave costs a whole
  // wave (config 2 on 8 GPUs: 1550 tiles = 10.47 waves of 148 -> 11). Its r tiles are then computed by a second launch of the
  // 64 x 64 kernel as 4 r sub-tiles written into the same packed tiles (gram_tile_of, TAIL mode): ceil(4 r / SMs) quarter-cost waves.
  int tail_r = 0;
  if (!narrow && !small && L.packed && L.tcol_begin < 0 && !getenv("PLB200_GRAM_NO_TAIL")) {
    const int64_t waves = mine / g.num_sms, r = mine % g.num_sms;
    if (waves >= 1 && waves <= 48 && r > 0 && 4 * r <= 3 * (int64_t)g.num_sms) {
      tail_r = (int)r;
      p.ntiles_total = L.part + (mine - r) * L.nparts;  // the main launch stops before this part's last r slots
    }
  }
  const int64_t main_tiles = mine - tail_r;
  const int grid = (int)(main_tiles < g.num_sms ? main_tiles : g.num_sms);
  // Dense results of the DotProduct / IMQ / LINEAR kernels leave through the TMA-store variant (smem-staged epilogue): 9-16 % faster when
  // the contraction is shallow and the kernel is bound by writing K, still 1-2 % at d = 300. Needs an output the TMA unit can address
  // (16-byte aligned, even leading dimension). PLB200_GRAM_TMAST=0/1 overrides the rule (experiments).
  // Not for RBF: its epilogue is bound by instruction ISSUE (8 FP64 + ~8 integer / LDS instructions per element, two warps per
  // sub-partition), and the staged form adds a shared-memory store and its address arithmetic per element: measured 3.70 vs 3.65 ms at
  // config 2 (profiles/r02b_epi_probe2.json).
  bool tmast = !narrow && !small && !L.packed && L.epi != EPI_RBF && !(reinterpret_cast<uintptr_t>(L.out) & 15u) && !(L.ldo & 1);
  if (const char* e = getenv("PLB200_GRAM_TMAST")) {
    if (e[0] == '0') tmast = false;
    else if (e[0] == '1') tmast = !narrow && !small && !L.packed && !(reinterpret_cast<uintptr_t>(L.out) & 15u) && !(L.ldo & 1);
  }
  CUtensorMap tmO = tmA;
  p.tma_out = 0;
  if (tmast) {
    PL_TRY(make_tmap_2d(&tmO, L.out, L.nB, L.nA, L.ldo, 16, 32));
    p.tma_out = 1;
  }
  if (L.epi != EPI_LINEAR || L.packed || L.symmetric) MAIN_KERNEL_BEGIN(st);
  cudaError_t le = cudaSuccess;
#define GRAM_LAUNCH(NARROW_, TMAST_, SMEM_, GRID_, PDL_, TA_, TB_, P_)                                                                        \
  switch (L.epi) {                                                                                                                            \
    case EPI_LINEAR: le = launch_ex(gram_kernel<EPI_LINEAR, NARROW_, TMAST_>, GRID_, GRAM_THREADS, SMEM_, st, PDL_, TA_, TB_, tmO, P_); break; \
    case EPI_DOT: le = launch_ex(gram_kernel<EPI_DOT, NARROW_, TMAST_>, GRID_, GRAM_THREADS, SMEM_, st, PDL_, TA_, TB_, tmO, P_); break;       \
    case EPI_RBF: le = launch_ex(gram_kernel<EPI_RBF, NARROW_, TMAST_>, GRID_, GRAM_THREADS, SMEM_, st, PDL_, TA_, TB_, tmO, P_); break;       \
    case EPI_IMQ: le = launch_ex(gram_kernel<EPI_IMQ, NARROW_, TMAST_>, GRID_, GRAM_THREADS, SMEM_, st, PDL_, TA_, TB_, tmO, P_); break;       \
    default: return set_err(PL_EINVAL, "unknown epilogue %d", L.epi);                                                                        \
  }
  if (narrow) {
    GRAM_LAUNCH(1, 0, GRAM_SMEM_BYTES + tab_smem, grid, pdl, tmA, tmB, p)
  } else if (small) {
    GRAM_LAUNCH(2, 0, GRAM_SMEM_BYTES + tab_smem, grid, pdl, tmA, tmB, p)
  } else if (tmast) {
    GRAM_LAUNCH(0, 1, GRAM_SMEM_BYTES_TMAST + tab_smem, grid, pdl, tmA, tmB, p)
  } else {
    GRAM_LAUNCH(0, 0, GRAM_SMEM_BYTES + tab_smem, grid, pdl, tmA, tmB, p)
  }
  CUDA_TRY(le);
  LAUNCH_CHECK();
#ifdef PLB_GRAM_TRACE
  if (p.trace) {
    std::vector<long long> h((size_t)g.num_sms * 8);
    CUDA_TRY(cudaMemcpyAsync(h.data(), p.trace, h.size() * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (FILE* f = fopen(getenv("PLB200_GRAM_TRACE"), "a")) {
      fprintf(f, "launch grid %d\n", grid);
      for (int b = 0; b < grid; ++b) {
        fprintf(f, "%d", b);
        for (int k = 0; k < 8; ++k) fprintf(f, " %lld", h[(size_t)b * 8 + k]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
#endif
  if (tail_r) {
    CUtensorMap tmA64, tmB64;
    PL_TRY(make_tmap_2d(&tmA64, L.A, L.d + L.b_shift, L.nA, L.ldA, GRAM_BK, 64));
    PL_TRY(make_tmap_2d(&tmB64, L.B, L.d + L.b_shift, L.nB, L.ldB, GRAM_BK, 64));
    GramParams pt = p;
    pt.T1 = (int)((L.nA + 63) / 64);
    pt.T2 = (int)((L.nB + 63) / 64);
    pt.tail_first = p.ntiles_total;  // first tail slot of this part
    pt.tail_count = tail_r;
    pt.tma_out = 0;
    const bool tpdl = !getenv("PLB200_NO_PDL");  // directly behind the main launch: its prologue overlaps the main kernel's last wave
    pt.pdl = tpdl ? 1 : 0;
    const int tgrid = (int)(4 * (int64_t)tail_r < g.num_sms ? 4 * tail_r : g.num_sms);
    GRAM_LAUNCH(2, 0, GRAM_SMEM_BYTES + tab_smem, tgrid, tpdl, tmA64, tmB64, pt)
    CUDA_TRY(le);
    LAUNCH_CHECK();
  }
#undef GRAM_LAUNCH
  if (L.epi != EPI_LINEAR || L.packed || L.symmetric) MAIN_KERNEL_END(st);
  return PL_OK;
}

template <int MODE>
int launch_prep(const double* X, int64_t n, int64_t d, int64_t ldx, const double* mean, const double* cdiag, double* Xc, double* Y, int64_t ldc,
                double* stat, cudaStream_t st) {
  if (n == 0) return PL_OK;
  const int64_t threads = n * 32;
  prep_rows_kernel<MODE><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(X, n, d, ldx, mean, cdiag, Xc, Y, ldc, stat);
  LAUNCH_CHECK();
  return PL_OK;
}

// prepared != NULL: run the preludes, fill *prepared with the main launch and return WITHOUT launching it (the host
// streaming path launches it several times over tile-column ranges).
int gram_dev_impl(int kernel_id, const double* dX1, int64_t n1, int64_t ldx1, const double* dX2, int64_t n2, int64_t ldx2, int64_t d, int alpha,
                  int cinv_kind, const double* dcinv, double ell, double beta, double* dOut, int64_t ldo, int fill, int part, int nparts, int packed,
                  cudaStream_t st, GramLaunch* prepared = nullptr) {
  PL_TRY(need_ready());
  const bool sym = (dX2 == nullptr);
  if (n1 < 0 || d <= 0 || (!sym && n2 < 0)) return set_err(PL_EINVAL, "bad shape n1=%lld n2=%lld d=%lld", (long long)n1, (long long)n2, (long long)d);
  if (!dX1 || !dOut) return set_err(PL_EINVAL, "null pointer");
  if (ldx1 < d || (!sym && ldx2 < d)) return set_err(PL_EINVAL, "leading dimension smaller than d");
  if (nparts < 1 || part < 0 || part >= nparts) return set_err(PL_EINVAL, "bad part %d of %d", part, nparts);
  if (sym) {
    n2 = n1;
    if (!packed && (fill < 1 || fill > 3)) return set_err(PL_EINVAL, "fill must be PL_FILL_RM_UPPER, PL_FILL_RM_LOWER or PL_FILL_FULL");
  }
  if (!packed && ldo < n2) return set_err(PL_EINVAL, "ldo smaller than the number of columns");
  if (n1 == 0 || n2 == 0) return PL_OK;

  const double *X1, *X2 = nullptr;
  int64_t l1, l2 = 0;
  PL_TRY(tma_operand(dX1, n1, d, ldx1, WS_PAD1, st, &X1, &l1));
  if (!sym) PL_TRY(tma_operand(dX2, n2, d, ldx2, WS_PAD2, st, &X2, &l2));

  GramLaunch L{};
  L.d = d;
  L.out = dOut;
  L.ldo = ldo;
  L.symmetric = sym;
  L.fill = fill;
  L.packed = packed;
  L.part = part;
  L.nparts = nparts;
  L.alpha = alpha;
  L.nA = n1;
  L.nB = n2;

  if (kernel_id == PL_KERNEL_LINEAR) {
    L.epi = EPI_LINEAR;
    L.A = X1; L.ldA = l1;
    L.B = sym ? X1 : X2; L.ldB = sym ? l1 : l2;
    L.same_operand = sym;
    if (prepared) { *prepared = L; return PL_OK; }
    return launch_gram(L, st);
  }
  if (kernel_id == PL_KERNEL_DOT) {
    void *s1, *s2 = nullptr;
    PL_TRY(ws_get(WS_STAT1, (size_t)n1 * 8, &s1));
    PL_TRY(launch_prep<PREP_DOT>(X1, n1, d, l1, nullptr, nullptr, nullptr, nullptr, 0, (double*)s1, st));
    if (!sym) {
      PL_TRY(ws_get(WS_STAT2, (size_t)n2 * 8, &s2));
      PL_TRY(launch_prep<PREP_DOT>(X2, n2, d, l2, nullptr, nullptr, nullptr, nullptr, 0, (double*)s2, st));
    }
    L.epi = EPI_DOT;
    L.pdl = 1;  // follows prep_rows: the main loop overlaps it, the epilogue waits for the row statistics
    L.A = X1; L.ldA = l1;
    L.B = sym ? X1 : X2; L.ldB = sym ? l1 : l2;
    L.ra = (double*)s1;
    L.rb = sym ? (double*)s1 : (double*)s2;
    L.same_operand = sym;
    if (prepared) { *prepared = L; return PL_OK; }
    return launch_gram(L, st);
  }
  if (kernel_id != PL_KERNEL_RBF && kernel_id != PL_KERNEL_IMQ) return set_err(PL_EINVAL, "unknown kernel id %d", kernel_id);
  if (!(ell != 0.0)) return set_err(PL_EINVAL, "length scale must be non-zero");
  if (kernel_id == PL_KERNEL_IMQ && !(beta > 0.0 && ell > 0.0)) return set_err(PL_EINVAL, "InverseMultiquadric needs 0 < c2 and 0 < ell (kernels.jl:143-146)");
  if (cinv_kind != PL_CINV_IDENTITY && cinv_kind != PL_CINV_DIAGONAL && cinv_kind != PL_CINV_FULL) return set_err(PL_EINVAL, "unknown cinv_kind %d", cinv_kind);
  if (cinv_kind != PL_CINV_IDENTITY && !dcinv) return set_err(PL_EINVAL, "cinv is NULL");

  // The distance is translation invariant: shift by the column mean of X1 so that the Gram form
  // q_i + q_j - 2 G_ij does not cancel against |mean|^2 (parity at 1e-10).
  void *mean, *xc1, *xc2 = nullptr, *y1 = nullptr, *y2 = nullptr, *s1, *s2 = nullptr;
  const int64_t ldc = (d + 1) & ~(int64_t)1;
  PL_TRY(ws_get(WS_MEAN, (size_t)ldc * 8, &mean));
  {
    // any shift near the centre of the data removes the cancellation: the mean of a strided sample of <= 512 rows, one small launch
    // (a full column-mean pass costs 26 us of the 530 us an 8-GPU shard of config 2 takes). Same sample on every rank -> same shift -> same bits.
    const int64_t stride = (n1 + 511) / 512;
    const int64_t ns = (n1 + stride - 1) / stride;
    sample_mean_kernel<<<(unsigned)((d + 31) / 32), 1024, 0, st>>>(X1, ns, l1 * stride, d, (double*)mean);
    LAUNCH_CHECK();
  }
  if (kernel_id == PL_KERNEL_RBF) {
    // RBF: operands in the augmented, pre-scaled layout (prelude.cuh rbf_aug_kernel): the accumulator of the Gram kernel is the
    // argument of 2^(./256) itself, the epilogue is exp2_tab and nothing else (gram.cuh)
    const int KT = (int)((d + GRAM_AUG_COLS + GRAM_BK - 1) / GRAM_BK);
    const int64_t lda = 16 * (int64_t)KT + 16;
    const double Lc = (double)EXP_TAB_N * 1.4426950408889634074;  // 256 log2(e)
    const double cs = sqrt(Lc) / fabs(ell);
    const double hb = beta > 0.0 ? 0.5 * Lc * log(beta) : 0.0;  // beta <= 0 cannot be folded into the exponent: epilogue multiplies (bmul)
    PL_TRY(ws_get(WS_XC1, (size_t)n1 * lda * 8, &xc1));
    if (!sym) PL_TRY(ws_get(WS_XC2, (size_t)n2 * lda * 8, &xc2));
    L.epi = EPI_RBF;
    L.pdl = 1;
    L.s = -0.5 / (ell * ell);
    L.beta = beta;
    L.d = 16 * (int64_t)KT;
    L.b_shift = 16;
    L.same_operand = 0;  // the two operands of a diagonal tile differ in their last k-step
    L.ldA = L.ldB = lda;
    auto aug = [&](int mode, const double* X, int64_t n, int64_t ldx, const double* cd, double* Xc, double* Y) -> int {
      if (n == 0) return PL_OK;
      const int blocks = (int)((n * 32 + 255) / 256);
      // an ordinary launch: as a programmatic dependent of sample_mean_kernel it measured no gain on plain launches and cost 30 us inside the
      // cached CUDA graph of a tile-packed shard with its tail launch (profiles/r02c notes)
      const bool apdl = getenv("PLB200_AUG_PDL") != nullptr;
      const double* mn = (const double*)mean;
      cudaError_t le;
      switch (mode) {
        case AUG_ID: le = launch_ex(rbf_aug_kernel<AUG_ID>, blocks, 256, 0, st, apdl, X, n, d, ldx, mn, cd, Xc, Y, KT, cs, hb); break;
        case AUG_DIAG: le = launch_ex(rbf_aug_kernel<AUG_DIAG>, blocks, 256, 0, st, apdl, X, n, d, ldx, mn, cd, Xc, Y, KT, cs, hb); break;
        case AUG_CENTER: le = launch_ex(rbf_aug_kernel<AUG_CENTER>, blocks, 256, 0, st, apdl, X, n, d, ldx, mn, cd, Xc, Y, KT, cs, hb); break;
        default: le = launch_ex(rbf_aug_kernel<AUG_FINISH>, blocks, 256, 0, st, apdl, X, n, d, ldx, mn, cd, Xc, Y, KT, cs, hb); break;
      }
      CUDA_TRY(le);
      LAUNCH_CHECK();
      return PL_OK;
    };
    if (cinv_kind == PL_CINV_IDENTITY) {
      PL_TRY(aug(AUG_ID, X1, n1, l1, nullptr, (double*)xc1, nullptr));
      if (!sym) PL_TRY(aug(AUG_ID, X2, n2, l2, nullptr, (double*)xc2, nullptr));
      L.A = (double*)xc1;
    } else {
      PL_TRY(ws_get(WS_Y1, (size_t)n1 * lda * 8, &y1));
      if (!sym) PL_TRY(ws_get(WS_Y2, (size_t)n2 * lda * 8, &y2));
      if (cinv_kind == PL_CINV_DIAGONAL) {
        const double* cd;
        int64_t ldcd;
        PL_TRY(tma_operand(dcinv, 1, d, ldc, WS_PADC, st, &cd, &ldcd));
        PL_TRY(aug(AUG_DIAG, X1, n1, l1, cd, (double*)xc1, (double*)y1));
        if (!sym) PL_TRY(aug(AUG_DIAG, X2, n2, l2, cd, (double*)xc2, (double*)y2));
      } else {
        // Y = Xc * Cinv through the same DMMA Gram engine (Cinv symmetric: rows of Cinv are its columns), then the row terms
        const double* cf;
        int64_t ldcf;
        PL_TRY(tma_operand(dcinv, d, d, d, WS_PADC, st, &cf, &ldcf));
        for (int side = 0; side < (sym ? 1 : 2); ++side) {
          const double* X = side ? X2 : X1;
          const int64_t n = side ? n2 : n1, lx = side ? l2 : l1;
          double* xc = (double*)(side ? xc2 : xc1);
          double* y = (double*)(side ? y2 : y1);
          PL_TRY(aug(AUG_CENTER, X, n, lx, nullptr, xc, nullptr));
          GramLaunch Ly{};
          Ly.epi = EPI_LINEAR;
          Ly.A = xc; Ly.nA = n; Ly.ldA = lda;
          Ly.B = cf; Ly.nB = d; Ly.ldB = ldcf;
          Ly.d = d;
          Ly.out = y; Ly.ldo = lda;
          Ly.nparts = 1;
          PL_TRY(launch_gram(Ly, st));
          PL_TRY(aug(AUG_FINISH, nullptr, n, 0, nullptr, xc, y));
        }
      }
      L.A = (double*)y1;
    }
    L.B = sym ? (double*)xc1 : (double*)xc2;
    if (prepared) { *prepared = L; return PL_OK; }
    return launch_gram(L, st);
  }
  // InverseMultiquadric: row statistics in the epilogue
  PL_TRY(ws_get(WS_XC1, (size_t)n1 * ldc * 8, &xc1));
  PL_TRY(ws_get(WS_STAT1, (size_t)n1 * 8, &s1));
  if (!sym) {
    PL_TRY(ws_get(WS_XC2, (size_t)n2 * ldc * 8, &xc2));
    PL_TRY(ws_get(WS_STAT2, (size_t)n2 * 8, &s2));
  }
  L.epi = EPI_IMQ;
  L.pdl = 1;  // follows prep_rows / rowdot: barrier set-up and descriptor prefetch overlap its tail
  L.s = 1.0 / (ell * ell);
  L.beta = beta;  // c2
  L.ra = (double*)s1;
  L.rb = sym ? (double*)s1 : (double*)s2;
  if (cinv_kind == PL_CINV_IDENTITY) {
    PL_TRY(launch_prep<PREP_DIST_ID>(X1, n1, d, l1, (double*)mean, nullptr, (double*)xc1, nullptr, ldc, (double*)s1, st));
    if (!sym) PL_TRY(launch_prep<PREP_DIST_ID>(X2, n2, d, l2, (double*)mean, nullptr, (double*)xc2, nullptr, ldc, (double*)s2, st));
    L.A = (double*)xc1; L.ldA = ldc;
    L.B = sym ? (double*)xc1 : (double*)xc2; L.ldB = ldc;
    L.same_operand = sym;
    if (prepared) { *prepared = L; return PL_OK; }
    return launch_gram(L, st);
  }
  PL_TRY(ws_get(WS_Y1, (size_t)n1 * ldc * 8, &y1));
  if (!sym) PL_TRY(ws_get(WS_Y2, (size_t)n2 * ldc * 8, &y2));
  if (cinv_kind == PL_CINV_DIAGONAL) {
    const double* cd;
    int64_t ldcd;
    PL_TRY(tma_operand(dcinv, 1, d, ldc, WS_PADC, st, &cd, &ldcd));
    PL_TRY(launch_prep<PREP_DIST_DIAG>(X1, n1, d, l1, (double*)mean, cd, (double*)xc1, (double*)y1, ldc, (double*)s1, st));
    if (!sym) PL_TRY(launch_prep<PREP_DIST_DIAG>(X2, n2, d, l2, (double*)mean, cd, (double*)xc2, (double*)y2, ldc, (double*)s2, st));
  } else {
    // Y = Xc * Cinv through the same DMMA Gram engine (Cinv symmetric: rows of Cinv are its columns)
    const double* cf;
    int64_t ldcf;
    PL_TRY(tma_operand(dcinv, d, d, d, WS_PADC, st, &cf, &ldcf));
    PL_TRY(launch_prep<PREP_CENTER_ONLY>(X1, n1, d, l1, (double*)mean, nullptr, (double*)xc1, nullptr, ldc, nullptr, st));
    GramLaunch Ly{};
    Ly.epi = EPI_LINEAR;
    Ly.A = (double*)xc1; Ly.nA = n1; Ly.ldA = ldc;
    Ly.B = cf; Ly.nB = d; Ly.ldB = ldcf;
    Ly.d = d;
    Ly.out = (double*)y1; Ly.ldo = ldc;
    Ly.nparts = 1;
    PL_TRY(launch_gram(Ly, st));
    rowdot_kernel<<<(unsigned)((n1 * 32 + 255) / 256), 256, 0, st>>>((double*)y1, (double*)xc1, n1, d, ldc, (double*)s1);
    LAUNCH_CHECK();
    if (!sym) {
      PL_TRY(launch_prep<PREP_CENTER_ONLY>(X2, n2, d, l2, (double*)mean, nullptr, (double*)xc2, nullptr, ldc, nullptr, st));
      Ly.A = (double*)xc2; Ly.nA = n2;
      Ly.out = (double*)y2;
      PL_TRY(launch_gram(Ly, st));
      rowdot_kernel<<<(unsigned)((n2 * 32 + 255) / 256), 256, 0, st>>>((double*)y2, (double*)xc2, n2, d, ldc, (double*)s2);
      LAUNCH_CHECK();
    }
  }
  L.A = (double*)y1; L.ldA = ldc;
  L.B = sym ? (double*)xc1 : (double*)xc2; L.ldB = ldc;
  L.same_operand = 0;
  if (prepared) { *prepared = L; return PL_OK; }
  return launch_gram(L, st);
}

// ---- SYRK schedule: which CTA runs which (work unit, row range) pieces ----
// One CTA is resident per SM. Work units are not equal (a diagonal unit costs ~1.06-1.19 x an off-diagonal one) and SMs do not all run
// at the same speed, so two schedules are built here as plain tables the kernel walks (syrk.cuh: SyrkPiece):
//   * "waves" (PLB200_SYRK_SCHED=waves): one piece per CTA, ~20 short CTAs per SM dealt out by the hardware as SMs free up (measured
//     with the trace hook: 5-6 long CTAs per SM left 9 % of the machine idle at the tail), each streaming >= 1024 rows;
//   * "streamk" (default; PLB200_SYRK_SCHED=sk<w> caps the rounds at w): ~148 x 4 CTAs of equal COST — a unit is cut into N0
//     (off-diagonal) or ceil(1.15 N0) (diagonal) long row chunks, cuts aligned across units, CTAs issued in row order, N0 chosen so
//     that the last round of CTAs is full. ~4 pipeline fills per SM instead
//     of 20, ~600 partials of 160 KB instead of ~3000, and the w-fold over-decomposition still lets the hardware even out SM speed
//     differences. Measured (profiles/r01_syrk_schedules.json): config 3 22.5 -> 22.2 ms, a per-GPU eighth of it 2.95 -> 2.82 ms,
//     a quarter of config 4 80.3 -> 78.2 ms; 4..8 rounds and cost 1.12..1.16 are flat.
// Either way the partial slots of a unit are contiguous and in row order, so the second-stage sum is deterministic.
struct SyrkSchedConfig {
  int mode;        // 0 waves, 1 streamk
  int waves;       // CTAs per SM
  double diag_cost;
  int min_rounds;  // streamk: at least this many rounds of CTAs (experiment knob PLB200_SYRK_MIN_ROUNDS; 0 = the model decides)
};

const SyrkSchedConfig& syrk_sched_config() {
  static const SyrkSchedConfig cfg = [] {
    SyrkSchedConfig c;
    c.mode = 1;
    c.waves = 0;
    c.diag_cost = 1.15;  // measured optimum 1.12-1.16 (flat); the DMMA count alone says 1.125
    const char* e = getenv("PLB200_SYRK_SCHED");
    if (e && !strncmp(e, "waves", 5)) c.mode = 0;
    else if (e && !strncmp(e, "sk", 2)) c.waves = atoi(e + 2);
    const char* w = getenv("PLB200_SYRK_WAVES");
    if (w && atoi(w) > 0) c.waves = atoi(w);
    if (c.waves <= 0) c.waves = 20;  // waves: CTAs per SM; streamk: the most rounds the chunk-count search may use
    const char* dc = getenv("PLB200_SYRK_DIAG_COST");
    if (dc && atof(dc) > 0.0) c.diag_cost = atof(dc);
    const char* mr = getenv("PLB200_SYRK_MIN_ROUNDS");
    c.min_rounds = mr ? atoi(mr) : 0;
    return c;
  }();
  return cfg;
}

void build_syrk_schedule(int64_t R, int T, int nunits, int num_sms, const SyrkSchedConfig& cfg, SyrkSchedule* out) {
  out->pieces.clear();
  out->cta_first.clear();
  out->unit_first.assign(nunits + 1, 0);
  const int64_t nst = (R + SYRK_BK - 1) / SYRK_BK;  // pipeline stages per unit
  if (nst == 0) {
    out->nctas = 1;
    out->nslots = 0;
    out->cta_first.assign(2, 0);
    return;
  }
  if (cfg.mode == 0) {
    int64_t s = ((int64_t)num_sms * cfg.waves + nunits - 1) / nunits;
    int64_t max_by_rows = R / 1024;
    if (max_by_rows < 1) max_by_rows = 1;
    if (s > max_by_rows) s = max_by_rows;
    if (s < 1) s = 1;
    int64_t sps = (nst + s - 1) / s;  // stages per split
    s = (nst + sps - 1) / sps;
    for (int64_t sp = 0; sp < s; ++sp)
      for (int u = 0; u < nunits; ++u) {  // unit fastest: the dispatch order of the (unit, split) grid
        const int64_t f = sp * sps;
        const int64_t n = f + sps <= nst ? sps : nst - f;
        out->cta_first.push_back((int)out->pieces.size());
        out->pieces.push_back(SyrkPiece{u, (int)f, (int)n, (int)(u * s + sp)});
      }
    out->cta_first.push_back((int)out->pieces.size());
    for (int u = 0; u <= nunits; ++u) out->unit_first[u] = (int)(u * s);
    out->nctas = (int)(s * nunits);
    out->nslots = (int)(s * nunits);
    return;
  }
  // streamk: every unit is cut into equal row chunks — N0 chunks for an off-diagonal unit, ceil(diag_cost * N0) for a diagonal one
  // (equal cost per CTA to within rounding, the diagonal CTAs never the longer ones), N0 the largest count that keeps the total at
  // or below 148 x waves CTAs. Units of the same kind get IDENTICAL cuts, and CTAs are issued in order of their first row: the CTAs
  // that run at the same time then read the same rows for different units, so a row range is fetched from HBM once and shared
  // through L2 (cutting the unit-major linearised work into equal-cost intervals instead left every unit streaming its own rows
  // alone: ncu showed 13.0 GB of DRAM reads for a 2.9 GB shard).
  const int noff = T * (T - 1) / 2, ndiag = nunits - noff;
  int64_t cap = nst / 64;  // a CTA should stream >= 64 stages (1024 rows)
  if (cap < 1) cap = 1;
  auto ndg = [&](int64_t n0) { return (int64_t)ceil(cfg.diag_cost * (double)n0 - 1e-9); };
  // N0 is chosen by a small model: CTAs of cost c run in ceil(nctas / SMs) rounds, so the efficiency is total cost / (SMs x rounds x c);
  // every extra wave costs ~0.15 % (one more pipeline fill and 160 KB partial per SM; measured: 4..8 waves flat, 20 waves 2 % slower).
  // With few units (the BASELINE shapes: 8 at K = 500, 32 at K = 1000) this lands on ~4 waves that fill the last round to > 99 %; with
  // more units than SMs it takes as many waves as the fill needs, up to cfg.waves.
  const double total_cost = (double)noff + cfg.diag_cost * (double)ndiag;
  int64_t N0 = noff ? 1 : 0, ND = 1;
  double best = -1.0;
  for (int64_t n = 1; n <= cap; ++n) {
    const int64_t nd = noff ? (ndg(n) < cap ? ndg(n) : cap) : n;
    const int64_t nctas = noff ? noff * n + ndiag * nd : ndiag * nd;
    const int64_t rounds = (nctas + num_sms - 1) / num_sms;
    if (rounds > cfg.waves && best >= 0.0) break;
    if (rounds < cfg.min_rounds) continue;
    const double c = noff ? std::max(1.0 / (double)n, cfg.diag_cost / (double)nd) : cfg.diag_cost / (double)nd;
    // fewer than 4 rounds leave the hardware too little to even out SM speed differences with (measured: 1 round -2.7 %, 2 rounds
    // -0.2..-1.1 %, 3 rounds -0.1..-0.9 % against 4)
    const double few = rounds == 1 ? 0.025 : (rounds == 2 ? 0.008 : (rounds == 3 ? 0.004 : 0.0));
    const double score = total_cost / ((double)num_sms * (double)rounds * c) - 0.0015 * (double)rounds - few;
    if (score > best + 1e-12) {
      best = score;
      N0 = noff ? n : 0;
      ND = nd;
    }
  }
  struct Cta {
    int64_t key;
    SyrkPiece pc;
  };
  std::vector<Cta> ctas;
  int slot = 0;
  for (int u = 0; u < nunits; ++u) {
    const int64_t n = u < noff ? N0 : ND;
    out->unit_first[u] = slot;
    for (int64_t j = 0; j < n; ++j) {
      const int64_t f = (nst * j) / n, e = (nst * (j + 1)) / n;
      if (e <= f) continue;
      ctas.push_back(Cta{f * nunits + u, SyrkPiece{u, (int)f, (int)(e - f), slot}});
      ++slot;
    }
  }
  out->unit_first[nunits] = slot;
  std::sort(ctas.begin(), ctas.end(), [](const Cta& x, const Cta& y) { return x.key < y.key; });
  for (const Cta& c : ctas) {
    out->cta_first.push_back((int)out->pieces.size());
    out->pieces.push_back(c.pc);
  }
  out->cta_first.push_back((int)out->pieces.size());
  out->nctas = (int)ctas.size();
  out->nslots = slot;
}


// drho_out (centred mode, optional): rho = sum_r (a_r - mean) of this shard, the by-product of the one-sided centring
int syrk_dev_impl(const double* dA, int64_t R, int64_t Kf, int64_t lda, const double* dw, int64_t ne, double w_e, double w_f, const double* db,
                  const double* dmean, double* dS, double* dvec, cudaStream_t st, double* drho_out = nullptr) {
  PL_TRY(need_ready());
  if (R < 0 || Kf <= 0 || lda < Kf) return set_err(PL_EINVAL, "bad shape R=%lld Kf=%lld lda=%lld", (long long)R, (long long)Kf, (long long)lda);
  if (!dA || !dS) return set_err(PL_EINVAL, "null pointer");
  if (ne < 0) ne = 0;
  if (ne > R) ne = R;
  if (dmean && dvec) return set_err(PL_EINVAL, "centred SYRK does not produce moment vectors");
  const double* A;
  int64_t ld;
  PL_TRY(tma_operand(dA, R, Kf, lda, WS_PAD1, st, &A, &ld));

  SyrkParams p;
  p.R = R;
  p.Kf = Kf;
  p.w = dw;
  p.ne = ne;
  p.w_e = w_e;
  p.w_f = w_f;
  p.b = db;
  p.mean = dmean;
  p.T = (int)((Kf + SYRK_BT - 1) / SYRK_BT);
  p.ntiles = syrk_num_units(p.T);
  if (g_sched.R != R || g_sched.T != p.T || g_sched.dev == nullptr || g_sched.dev != g.ws[WS_SYRK_SCHED]) {
    build_syrk_schedule(R, p.T, p.ntiles, g.num_sms, syrk_sched_config(), &g_sched.sch);
    const SyrkSchedule& sc = g_sched.sch;
    const size_t n1 = sc.cta_first.size(), n2 = sc.unit_first.size(), n3 = sc.pieces.size() * 4;
    void* dsch;
    PL_TRY(ws_get(WS_SYRK_SCHED, (n1 + n2 + n3 + 4) * sizeof(int), &dsch));
    int* di = (int*)dsch;
    // pageable sources: the copies are staged before the calls return, so the vectors may change afterwards
    CUDA_TRY(cudaMemcpyAsync(di, sc.cta_first.data(), n1 * sizeof(int), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(di + n1, sc.unit_first.data(), n2 * sizeof(int), cudaMemcpyHostToDevice, st));
    if (n3) CUDA_TRY(cudaMemcpyAsync(di + ((n1 + n2 + 3) & ~(size_t)3), sc.pieces.data(), n3 * sizeof(int), cudaMemcpyHostToDevice, st));
    if (!g_sched.ready) CUDA_TRY(cudaEventCreateWithFlags(&g_sched.ready, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(g_sched.ready, st));
    g_sched.up_stream = st;
    g_sched.R = R;
    g_sched.T = p.T;
    g_sched.dev = dsch;
  } else if (st != g_sched.up_stream) {
    CUDA_TRY(cudaStreamWaitEvent(st, g_sched.ready, 0));
  }
  const SyrkSchedule& sc = g_sched.sch;
  {
    const size_t n1 = sc.cta_first.size(), n2 = sc.unit_first.size();
    const int* di = (const int*)g_sched.dev;
    p.cta_first = di;
    p.pieces = (const SyrkPiece*)(di + ((n1 + n2 + 3) & ~(size_t)3));  // 16-byte aligned
  }
  const int* d_unit_first = (const int*)g_sched.dev + sc.cta_first.size();
  void* wst;
  PL_TRY(ws_get(WS_SYRK_TILES, (size_t)(sc.nslots > 0 ? sc.nslots : 1) * SYRK_TILE_ELEMS * 8, &wst));
  void* rho = nullptr;
  if (dmean) PL_TRY(ws_get(WS_SYRK_RHO, (size_t)Kf * 8, &rho));
  p.ws_tiles = (double*)wst;

  // developer hook: PLB200_SYRK_TRACE=<file> dumps per-CTA timing (start, SM, end time of each consumer warp) of every SYRK launch
  const char* trace_path = getenv("PLB200_SYRK_TRACE");
  p.trace = nullptr;
  if (trace_path && *trace_path) {
    void* tr;
    PL_TRY(ws_get(WS_TRACE, (size_t)sc.nctas * 10 * 8, &tr));
    CUDA_TRY(cudaMemsetAsync(tr, 0, (size_t)sc.nctas * 10 * 8, st));
    p.trace = (long long*)tr;
  }

  CUtensorMap tm, tmx;
  // R == 0: the kernel loads nothing (KT = 0) but the descriptor still needs a non-zero extent
  PL_TRY(make_tmap_2d(&tm, A, Kf, R > 0 ? R : 1, ld, 16, SYRK_BK));
  // Moment vectors: ride along as two extra columns of the last (padded) column block whenever it has room — the box holding
  // column Kf is then loaded from an auxiliary R x 16 tensor (syrk_aux_box_kernel). Otherwise (Kf % 16 == 15 or Kf % 128 == 0):
  // a weighted column-sum pass after the SYRK.
  const bool aug = (dvec || dmean) && R > 0 && (Kf % 128) != 0 && (Kf % 16) <= 14 && !getenv("PLB200_SYRK_NO_AUG");
  p.aux_box = -1;
  tmx = tm;
  if (aug) {
    void* aux;
    PL_TRY(ws_get(WS_SYRK_AUX, (size_t)R * 16 * 8, &aux));
    syrk_aux_box_kernel<<<(unsigned)((R * 16 + 255) / 256), 256, 0, st>>>(A, R, Kf, ld, db, ne, dmean ? 1 : 0, (double*)aux);
    LAUNCH_CHECK();
    PL_TRY(make_tmap_2d(&tmx, (const double*)aux, 16, R, 16, 16, SYRK_BK));
    p.aux_box = (int)(Kf / 16);
  }
  dim3 grid((unsigned)sc.nctas);
  MAIN_KERNEL_BEGIN(st);
  if (dmean)
    syrk_kernel<true><<<grid, SYRK_THREADS, SYRK_SMEM_BYTES, st>>>(tm, tmx, p);
  else
    syrk_kernel<false><<<grid, SYRK_THREADS, SYRK_SMEM_BYTES, st>>>(tm, tmx, p);
  LAUNCH_CHECK();
  MAIN_KERNEL_END(st);
  if (p.trace) {
    const size_t cnt = (size_t)sc.nctas * 10;
    long long* h = (long long*)malloc(cnt * 8);
    if (h) {
      CUDA_TRY(cudaMemcpyAsync(h, p.trace, cnt * 8, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      FILE* f = fopen(trace_path, "w");
      if (f) {
        fprintf(f, "# units %d ctas %d T %d R %lld Kf %lld slots %d\n", p.ntiles, sc.nctas, p.T, (long long)R, (long long)Kf, sc.nslots);
        for (int q = 0; q < sc.nctas; ++q) {  // first unit of the CTA, CTA index, start, SM, end of the 8 consumer warps, stages
          const long long* e = h + (size_t)q * 10;
          const int pc0 = sc.cta_first[q], pc1 = sc.cta_first[q + 1];
          long long stages = 0;
          for (int pc = pc0; pc < pc1; ++pc) stages += sc.pieces[pc].nstages;
          fprintf(f, "%d %d %lld %lld", pc1 > pc0 ? sc.pieces[pc0].unit : -1, q, e[0], e[1]);
          for (int w = 0; w < 8; ++w) fprintf(f, " %lld", e[2 + w]);
          fprintf(f, " %lld %d\n", stages, pc1 - pc0);
        }
        fclose(f);
      }
      free(h);
    }
  }
  double* augvec = aug ? dvec : nullptr;
  if (dmean && aug) {  // rho arrives as the augmented column: reduce S' first, then subtract rho_i m_j
    void* tmp;
    PL_TRY(ws_get(WS_SYRK_VEC, (size_t)(2 * Kf + 2) * 8, &tmp));
    augvec = (double*)tmp;
  } else if (dmean) {
    // rho_i = sum_r (a_ri - m_i) closes the one-sided centring (S_ij = S'_ij - rho_i m_j): one HBM-rate pass, differences first
    PL_TRY(colsum(A, R, Kf, ld, 0.0, (double*)rho, st, dmean));
  }
  syrk_reduce_kernel<<<dim3((unsigned)p.ntiles, SYRK_TILE_ELEMS / 256), 256, 0, st>>>(p.ws_tiles, p.T, d_unit_first, Kf, dS, (dmean && !aug) ? dmean : nullptr,
                                                                                   (const double*)rho, augvec);
  LAUNCH_CHECK();
  if (dmean && aug) {
    int blocks = (int)((Kf * Kf + 255) / 256);
    if (blocks > g.num_sms * 8) blocks = g.num_sms * 8;
    syrk_center_fix_kernel<<<blocks, 256, 0, st>>>(dS, Kf, augvec, dmean);
    LAUNCH_CHECK();
  }
  if (dmean && drho_out) CUDA_TRY(cudaMemcpyAsync(drho_out, aug ? (const void*)augvec : (const void*)rho, (size_t)Kf * 8, cudaMemcpyDeviceToDevice, st));
  if (dvec && !aug) {
    PL_TRY(moment_vectors(A, R, Kf, ld, dw, ne, w_e, w_f, db, dvec, dvec + Kf, st));
    syrk_scalar_kernel<<<1, 1024, 0, st>>>(dw, db, ne, w_e, dvec + 2 * Kf);
    LAUNCH_CHECK();
  }
  return PL_OK;
}

int finish_dev_impl(const double* dS, const double* dvec, int64_t Kf, int ic, double lambda, double* dAtWA, double* dAtWb, cudaStream_t st) {
  PL_TRY(need_ready());
  if (Kf <= 0 || !dS || !dAtWA) return set_err(PL_EINVAL, "bad arguments");
  if ((ic || dAtWb) && !dvec) return set_err(PL_EINVAL, "intercept / AtWb need the moment vectors");
  ic = ic ? 1 : 0;
  const int64_t n = Kf + ic;
  int blocks = (int)((n * n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  normal_eq_finish_kernel<<<blocks, 256, 0, st>>>(dS, dvec, Kf, ic, lambda, dAtWA, dAtWb);
  LAUNCH_CHECK();
  return PL_OK;
}

// host <-> device helpers for the host-pointer entry points
int h2d_matrix(const double* h, int64_t rows, int64_t cols, int64_t ldh, int slot, double** dptr, int64_t* ldd) {
  const int64_t ldp = (cols + 1) & ~(int64_t)1;
  void* buf = nullptr;
  PL_TRY(ws_get(slot, (size_t)(rows > 0 ? rows : 1) * ldp * 8, &buf));
  if (rows > 0) {
    if (staging_for(h, (size_t)rows * cols * 8))  // pageable source: our own pinned staging ring + host threads
      CUDA_TRY(g_stager.h2d_2d(buf, (size_t)ldp * 8, h, (size_t)ldh * 8, (size_t)cols * 8, (size_t)rows, g.stream));
    else if (ldp == ldh && ldh == cols)  // dense on both sides: one linear copy (faster than the pitched path)
      CUDA_TRY(cudaMemcpyAsync(buf, h, (size_t)rows * cols * 8, cudaMemcpyHostToDevice, g.stream));
    else
      CUDA_TRY(cudaMemcpy2DAsync(buf, (size_t)ldp * 8, h, (size_t)ldh * 8, (size_t)cols * 8, (size_t)rows, cudaMemcpyHostToDevice, g.stream));
  }
  *dptr = (double*)buf;
  *ldd = ldp;
  return PL_OK;
}
int h2d_vector(const double* h, int64_t count, int slot, double** dptr) {
  void* buf = nullptr;
  PL_TRY(ws_get(slot, (size_t)(count > 0 ? count : 1) * 8, &buf));
  if (count > 0) CUDA_TRY(cudaMemcpyAsync(buf, h, (size_t)count * 8, cudaMemcpyHostToDevice, g.stream));
  *dptr = (double*)buf;
  return PL_OK;
}

int finish_timing() {
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  float a = 0, b = 0, c = 0;
  CUDA_TRY(cudaEventElapsedTime(&a, g.ev[0], g.ev[1]));
  CUDA_TRY(cudaEventElapsedTime(&b, g.ev[1], g.ev[2]));
  CUDA_TRY(cudaEventElapsedTime(&c, g.ev[2], g.ev[3]));
  g.t_h2d = a;
  g.t_kernel = b;
  g.t_d2h = c;
  g.t_host = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g.t_entry).count();
  return PL_OK;
}


int global_feature_dev_impl(const double* dL, int64_t ldl, const int64_t* doff, int64_t N, int64_t d, int mean, double* dX, int64_t ldx,
                            int64_t total_rows, cudaStream_t st) {
  PL_TRY(need_ready());
  if (N < 0 || d <= 0 || ldl < d || ldx < d || !dL || !doff || !dX) return set_err(PL_EINVAL, "bad arguments");
  if (N == 0) return PL_OK;
  const double* L;
  int64_t ld;
  PL_TRY(tma_operand(dL, total_rows, d, ldl, WS_PAD1, st, &L, &ld));  // 128-bit loads need 16-byte alignment and an even ld
  int G = 1;
  while (G < 32 && 2 * G < d) G <<= 1;
  const int64_t nchunks = (d + 2 * G - 1) / (2 * G);
  const int64_t threads = N * nchunks * G;
  segment_reduce_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(L, ld, doff, N, d, G, mean, dX, ldx);
  LAUNCH_CHECK();
  return PL_OK;
}

// offsets must be non-decreasing with non-empty segments (sum / mean of an empty collection throws in the reference)
int check_offsets(const int64_t* off, int64_t N) {
  if (off[0] != 0) return set_err(PL_EINVAL, "offsets[0] must be 0");
  for (int64_t i = 0; i < N; ++i)
    if (off[i + 1] <= off[i]) return set_err(PL_EINVAL, "configuration %lld has no local descriptors (offsets must increase strictly)", (long long)i);
  return PL_OK;
}

// First half: leaves the UNSCALED sums of these rows on the device as one bucket [S (Kf x Kf) ; vec = v, c, s (2 Kf + 2)] — what a
// multi-GPU caller all-reduces — and returns after enqueueing (g.ev[0], g.ev[1] recorded around the vector uploads).
int normal_eq_stream_partial(const double* const* blocks, const int64_t* rows, const int64_t* lds, int64_t nblocks, int64_t R, int64_t Kf,
                             const double* w, int64_t ne, double w_e, double w_f, const double* b, bool want_vec, int64_t chunk_rows,
                             double** bucket_out) {
  const int64_t ldp = (Kf + 1) & ~(int64_t)1;
  if (chunk_rows > R) chunk_rows = R;
  if (chunk_rows < 1) chunk_rows = 1;
  void *buf[2], *dS, *dvec = nullptr, *dSacc;
  double *dVacc, *dw = nullptr, *db = nullptr;
  PL_TRY(ws_get(WS_NE_BUF0, (size_t)chunk_rows * ldp * 8, &buf[0]));
  PL_TRY(ws_get(WS_NE_BUF1, (size_t)(R > chunk_rows ? chunk_rows : 1) * ldp * 8, &buf[1]));
  PL_TRY(ws_get(WS_H_S, (size_t)Kf * Kf * 8, &dS));
  {
    double* bk;
    PL_TRY(bucket_get((size_t)(Kf * Kf + 2 * Kf + 2), WS_NE_SACC, &bk));
    dSacc = bk;
  }
  dVacc = (double*)dSacc + Kf * Kf;
  if (want_vec) PL_TRY(ws_get(WS_H_VEC, (size_t)(2 * Kf + 2) * 8, &dvec));
  CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
  if (w) PL_TRY(h2d_vector(w, R, WS_H_IN2, &dw));
  if (b) PL_TRY(h2d_vector(b, R, WS_H_IN3, &db));
  CUDA_TRY(cudaMemsetAsync(dSacc, 0, (size_t)(Kf * Kf + 2 * Kf + 2) * 8, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
  // the copy stream must not overwrite buffers an earlier call's kernels may still be reading
  CUDA_TRY(cudaEventRecord(g.chunk_ev[4], g.stream));
  CUDA_TRY(cudaStreamWaitEvent(g.copy_stream, g.chunk_ev[4], 0));
  int64_t blk = 0, blk_row = 0;  // cursor into the block list
  int64_t r0 = 0;
  const bool stage_blocks = !getenv("PLB200_NO_STAGING") && (double)R * (double)Kf * 8.0 >= 4.0 * 1048576.0;
  bool staged = false;
  for (int64_t c = 0; r0 < R; ++c) {
    const int bi = (int)(c & 1);
    const int64_t nr = R - r0 < chunk_rows ? R - r0 : chunk_rows;
    if (c >= 2) CUDA_TRY(cudaStreamWaitEvent(g.copy_stream, g.chunk_ev[2 + bi], 0));  // SYRK of chunk c-2 has released the buffer
    int64_t filled = 0;
    while (filled < nr) {
      while (rows[blk] == blk_row) {
        ++blk;
        blk_row = 0;
      }
      int64_t take = rows[blk] - blk_row;
      if (take > nr - filled) take = nr - filled;
      const double* src = blocks[blk] + blk_row * lds[blk];
      double* dst = (double*)buf[bi] + filled * ldp;
      if (stage_blocks && HostStager::is_pageable(src)) {
        CUDA_TRY(g_stager.h2d_2d(dst, (size_t)ldp * 8, src, (size_t)lds[blk] * 8, (size_t)Kf * 8, (size_t)take, g.copy_stream, false));
        staged = true;
      } else if (lds[blk] == Kf && ldp == Kf)
        CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)take * Kf * 8, cudaMemcpyHostToDevice, g.copy_stream));
      else
        CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)ldp * 8, src, (size_t)lds[blk] * 8, (size_t)Kf * 8, (size_t)take, cudaMemcpyHostToDevice, g.copy_stream));
      filled += take;
      blk_row += take;
    }
    if (staged) CUDA_TRY(g_stager.finish());  // every DMA of this chunk has been enqueued on the copy stream
    CUDA_TRY(cudaEventRecord(g.chunk_ev[bi], g.copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(g.stream, g.chunk_ev[bi], 0));
    int64_t ne_c = ne - r0;
    ne_c = ne_c < 0 ? 0 : (ne_c > nr ? nr : ne_c);
    PL_TRY(syrk_dev_impl((const double*)buf[bi], nr, Kf, ldp, dw ? dw + r0 : nullptr, ne_c, w_e, w_f, db ? db + r0 : nullptr, nullptr, (double*)dS,
                         (double*)dvec, g.stream));
    accumulate_kernel<<<g.num_sms * 4, 256, 0, g.stream>>>((const double*)dS, Kf * Kf, (double*)dSacc);
    LAUNCH_CHECK();
    if (want_vec) {
      accumulate_kernel<<<8, 256, 0, g.stream>>>((const double*)dvec, 2 * Kf + 2, dVacc);
      LAUNCH_CHECK();
    }
    CUDA_TRY(cudaEventRecord(g.chunk_ev[2 + bi], g.stream));
    r0 += nr;
  }
  *bucket_out = (double*)dSacc;
  return PL_OK;
}

// Second half on the (possibly all-reduced) bucket: finish kernel, D2H, timing.
int normal_eq_bucket_finish(const double* bucket, int64_t Kf, bool want_vec, int ic, double lambda, double* AtWA, double* AtWb) {
  const int64_t n = Kf + ic;
  void *dM, *dv = nullptr;
  PL_TRY(ws_get(WS_H_OUT1, (size_t)n * n * 8, &dM));
  if (AtWb) PL_TRY(ws_get(WS_H_OUT2, (size_t)n * 8, &dv));
  PL_TRY(finish_dev_impl(bucket, want_vec ? bucket + Kf * Kf : nullptr, Kf, ic, lambda, (double*)dM, (double*)dv, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
  CUDA_TRY(cudaMemcpyAsync(AtWA, dM, (size_t)n * n * 8, cudaMemcpyDeviceToHost, g.stream));
  if (AtWb) CUDA_TRY(cudaMemcpyAsync(AtWb, dv, (size_t)n * 8, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[3], g.stream));
  return finish_timing();
}

// Streamed weighted normal equations over a list of host row blocks: chunks of rows are copied on the copy stream into two
// alternating device buffers while the SYRK of the previous chunk runs on the compute stream; the unscaled chunk partials are
// accumulated in chunk order (deterministic), then finished exactly like the resident path.
int normal_eq_stream(const double* const* blocks, const int64_t* rows, const int64_t* lds, int64_t nblocks, int64_t R, int64_t Kf, const double* w,
                     int64_t ne, double w_e, double w_f, const double* b, int ic, double lambda, double* AtWA, double* AtWb, int64_t chunk_rows) {
  const bool want_vec = ic || b;
  double* bucket = nullptr;
  PL_TRY(normal_eq_stream_partial(blocks, rows, lds, nblocks, R, Kf, w, ne, w_e, w_f, b, want_vec, chunk_rows, &bucket));
  return normal_eq_bucket_finish(bucket, Kf, want_vec, ic, lambda, AtWA, AtWb);
}


// ------------------------------------------------ symmetric eigendecomposition (row f1) ------------------------------------------------
// The dense symmetric eigensolver is a LIBRARY call, as it is in the reference (LAPACK syevr behind Julia's eigen(Symmetric), reached through
// Determinantal.jl's EllEnsemble and dimension_reduction.jl:13): cuSOLVER's Xsyevd, resolved with dlopen so that libplb200.so has no link-time
// dependency beyond the CUDA runtime. What is ours is everything around it: the kernel matrix never leaves the device, and the L-ensemble
// quantities that need all of U (inclusion probabilities, the selected eigenvectors) are computed / gathered there.
struct Cusolver {
  void* dl = nullptr;
  cusolverDnHandle_t h = nullptr;
  cusolverDnParams_t params = nullptr;
  cusolverStatus_t (*Create)(cusolverDnHandle_t*) = nullptr;
  cusolverStatus_t (*Destroy)(cusolverDnHandle_t) = nullptr;
  cusolverStatus_t (*SetStream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
  cusolverStatus_t (*CreateParams)(cusolverDnParams_t*) = nullptr;
  cusolverStatus_t (*DestroyParams)(cusolverDnParams_t) = nullptr;
  cusolverStatus_t (*XsyevdBufferSize)(cusolverDnHandle_t, cusolverDnParams_t, cusolverEigMode_t, cublasFillMode_t, int64_t, cudaDataType, const void*,
                                       int64_t, cudaDataType, const void*, cudaDataType, size_t*, size_t*) = nullptr;
  cusolverStatus_t (*Xsyevd)(cusolverDnHandle_t, cusolverDnParams_t, cusolverEigMode_t, cublasFillMode_t, int64_t, cudaDataType, void*, int64_t,
                             cudaDataType, void*, cudaDataType, void*, size_t, void*, size_t, int*) = nullptr;
} cs;

int cusolver_ready() {
  if (cs.h) return PL_OK;
  const char* env = getenv("PLB200_CUSOLVER");
  const char* cands[] = {env, "libcusolver.so.11", "/usr/local/cuda/lib64/libcusolver.so.11", "libcusolver.so", "libcusolver.so.12"};
  for (const char* c : cands) {
    if (!c || !*c) continue;
    cs.dl = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (cs.dl) break;
  }
  if (!cs.dl) return set_err(PL_ECUDA, "cannot load libcusolver (%s); set PLB200_CUSOLVER to its path", dlerror());
#define CS_SYM(field, name)                                                         \
  do {                                                                              \
    *(void**)(&cs.field) = dlsym(cs.dl, name);                                      \
    if (!cs.field) return set_err(PL_ECUDA, "libcusolver does not export %s", name); \
  } while (0)
  CS_SYM(Create, "cusolverDnCreate");
  CS_SYM(Destroy, "cusolverDnDestroy");
  CS_SYM(SetStream, "cusolverDnSetStream");
  CS_SYM(CreateParams, "cusolverDnCreateParams");
  CS_SYM(DestroyParams, "cusolverDnDestroyParams");
  CS_SYM(XsyevdBufferSize, "cusolverDnXsyevd_bufferSize");
  CS_SYM(Xsyevd, "cusolverDnXsyevd");
#undef CS_SYM
  cusolverStatus_t r = cs.Create(&cs.h);
  if (r != CUSOLVER_STATUS_SUCCESS) {
    cs.h = nullptr;
    return set_err(PL_ECUDA, "cusolverDnCreate failed with status %d", (int)r);
  }
  r = cs.CreateParams(&cs.params);
  if (r != CUSOLVER_STATUS_SUCCESS) return set_err(PL_ECUDA, "cusolverDnCreateParams failed with status %d", (int)r);
  return PL_OK;
}

// ---- the library-free path (eigh.cuh): Householder tridiagonalisation in one cooperative kernel, divide and conquer with the eigenvector
// updates on the DMMA Gram engine, back-transformation. Chosen by eigh_dev_impl for n <= PLB200_EIGH_OWN_MAX (default 2048; the matrix and
// its three work matrices live in L2, up to n = 1184 the matrix itself in registers); PLB200_EIGH=cusolver / own forces one path.
struct EighTree {
  int n = -1;
  std::vector<int2> leaves;
  std::vector<std::vector<EighMerge>> levels;  // by height: children are complete before their parent
  int kmax = 0;
};
EighTree g_eigh_tree;

int eigh_build_node(EighTree& t, int lo, int hi) {  // returns the height
  if (hi - lo <= EIGH_LEAF) {
    t.leaves.push_back(make_int2(lo, hi));
    return 0;
  }
  int m = (lo + hi) / 2;
  m += (m - lo) & 1;  // even child offsets: the blocks stay 16-byte aligned for the TMA operands of the Gram engine
  const int hl = eigh_build_node(t, lo, m), hr = eigh_build_node(t, m, hi);
  const int h = 1 + (hl > hr ? hl : hr);
  if ((int)t.levels.size() < h) t.levels.resize(h);
  t.levels[h - 1].push_back(EighMerge{lo, m, hi});
  if (hi - lo > t.kmax) t.kmax = hi - lo;
  return h;
}

int eigh_own_impl(double* dA, int64_t n64, int64_t lda, int fill, int want_vectors, double* dW, cudaStream_t st) {
  const int n = (int)n64;
  if (n > EIGH_MAX_N) return set_err(PL_EINVAL, "the built-in eigensolver takes n <= %d", EIGH_MAX_N);
  EighTree& T = g_eigh_tree;
  if (T.n != n) {
    T = EighTree();
    T.n = n;
    eigh_build_node(T, 0, n);
  }
  const int64_t ldq = (n + 3) & ~(int64_t)3;
  const int lower = (fill & PL_FILL_RM_LOWER) ? 1 : 0;
  void *pS, *pQA, *pQB, *pVT, *pSmall, *pTab;
  PL_TRY(ws_get(WS_EIGH_S, (size_t)n * ldq * 8, &pS));
  PL_TRY(ws_get(WS_EIGH_QA, (size_t)n * ldq * 8, &pQA));
  PL_TRY(ws_get(WS_EIGH_QB, (size_t)n * ldq * 8, &pQB));
  PL_TRY(ws_get(WS_EIGH_VT, (size_t)n * ldq * 8, &pVT));
  // small arrays (doubles): d e tau p rowbuf(2) lam dl wz what pd(256) rot(4n) scal(8) bar(8) ; ints: perm order nrot Kp ; rho per merge
  const size_t np8 = ((size_t)n + 7) & ~(size_t)7;
  const size_t nmerge_max = np8;
  const size_t small_doubles = np8 * 11 + 512 + 4 * np8 + 8 + EIGH_BAR_WORDS + nmerge_max + 8 * np8 + 64;
  const size_t small_ints = np8 * 2 + 2 * nmerge_max;
  PL_TRY(ws_get(WS_EIGH_SMALL, small_doubles * 8 + small_ints * 4, &pSmall));
  double* sd = (double*)pSmall;
  double *d_d = sd, *d_e = sd + np8, *d_tau = sd + 2 * np8, *d_p = sd + 3 * np8 /* 2 n */, *d_rowbuf = sd + 5 * np8 /* 2 n */,
         *d_lam = sd + 7 * np8, *d_dl = sd + 8 * np8, *d_wz = sd + 9 * np8, *d_what = sd + 10 * np8, *d_pd = sd + 11 * np8 /* 2 x 256 */;
  double4* d_rot = (double4*)(d_pd + 512);
  double* d_scal = (double*)(d_rot + np8);
  unsigned long long* d_bar = (unsigned long long*)(d_scal + 8);
  double* d_rho = (double*)(d_bar + EIGH_BAR_WORDS);
  double* d_Tf = d_rho + nmerge_max;  // R x R per block of R reflectors: <= 8 n + 64 doubles
  int* si = (int*)(d_Tf + 8 * np8 + 64);
  int *d_perm = si, *d_order = si + np8, *d_nrot = si + 2 * np8, *d_Kp = si + 2 * np8 + nmerge_max;

  // tables: leaves, then per level: merges, rows, row2merge
  std::vector<int> tab;
  auto push = [&](const void* src, size_t nints) {
    const size_t at = tab.size();
    tab.resize(at + ((nints + 3) & ~(size_t)3));
    memcpy(tab.data() + at, src, nints * 4);
    return at;
  };
  const size_t at_leaves = push(T.leaves.data(), T.leaves.size() * 2);
  struct LevelAt { size_t merges, rows, r2m; int nrows, nm, kmax; };
  std::vector<LevelAt> lat;
  for (auto& lv : T.levels) {
    std::vector<int> rows, r2m;
    int km = 0;
    for (size_t mi = 0; mi < lv.size(); ++mi) {
      for (int i = lv[mi].lo; i < lv[mi].hi; ++i) { rows.push_back(i); r2m.push_back((int)mi); }
      km = std::max(km, lv[mi].hi - lv[mi].lo);
    }
    LevelAt a;
    a.merges = push(lv.data(), lv.size() * 3);
    a.rows = push(rows.data(), rows.size());
    a.r2m = push(r2m.data(), r2m.size());
    a.nrows = (int)rows.size(); a.nm = (int)lv.size(); a.kmax = km;
    lat.push_back(a);
  }
  PL_TRY(ws_get(WS_EIGH_TAB, tab.size() * 4, &pTab));
  const int* dtab = (const int*)pTab;
  double *S = (double*)pS, *QA = (double*)pQA, *QB = (double*)pQB, *VT = (double*)pVT;
  CUDA_TRY(cudaMemsetAsync(d_bar, 0, EIGH_BAR_WORDS * 8, st));
  CUDA_TRY(cudaMemsetAsync(d_pd, 0, 512 * 8, st));
  CUDA_TRY(cudaMemsetAsync(QA, 0, (size_t)n * ldq * 8, st));
  unsigned long long* d_max = d_bar + EIGH_BAR_GO + 2;  // a free word of the barrier block: the max-abs accumulator
  int blocks = (int)(((int64_t)n * n + 255) / 256);
  if (blocks > g.num_sms * 8) blocks = g.num_sms * 8;
  eigh_maxabs_kernel<<<blocks, 256, 0, st>>>(dA, n, lda, lower, d_max);
  LAUNCH_CHECK();
  {
    // one synchronisation: the tables are a local vector, and a matrix with Inf / NaN entries is refused like LAPACK's chkfinite does
    // for eigen(Symmetric) in the reference (NaN compares greater than everything in the bit-pattern maximum)
    double mx = 0.0;
    CUDA_TRY(cudaMemcpyAsync(pTab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(&mx, d_max, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (!(mx <= 1.79769313486231570e308)) return set_err(PL_EINVAL, "eigensolver: the matrix contains Infs or NaNs");
  }
  eigh_symmetrize_kernel<<<blocks, 256, 0, st>>>(dA, n, lda, lower, d_max, S, ldq, d_rowbuf, d_scal);
  LAUNCH_CHECK();
  // 1. tridiagonalisation
  {
    SytrdParams P;
    P.S = S; P.lds = ldq; P.n = n; P.d = d_d; P.e = d_e; P.tau = d_tau; P.p = d_p; P.pd = d_pd; P.rowbuf = d_rowbuf; P.bar = d_bar;
    P.prof = getenv("PLB200_EIGH_PROF") ? (long long*)d_rho : nullptr;  // over d_rho/d_Tf: written at the end of sytrd, read back below
    static int attr_dev = -1;  // function attributes are per device: set them again when the library is re-initialised on another one
    if (attr_dev != g.device) {
      CUDA_TRY(cudaFuncSetAttribute(eigh_sytrd_l2_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (4 * 2048 + 16 * 1024) * 8));
      CUDA_TRY(cudaFuncSetAttribute(eigh_sytrd_l2_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 2048 * 8));
      CUDA_TRY(cudaFuncSetAttribute(eigh_sytrd_l2_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 4096 * 8));
      CUDA_TRY(cudaFuncSetAttribute(dc_deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * EIGH_MAX_N));
      CUDA_TRY(cudaFuncSetAttribute(dc_small_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 32 * (EIGH_SMALL_GEMM_MAX | 1) * 8));
      CUDA_TRY(cudaFuncSetAttribute(eigh_backtransform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      CUDA_TRY(cudaFuncSetAttribute(eigh_backtransform_wy_kernel<8, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8 * 8 * 32 * 8));
      CUDA_TRY(cudaFuncSetAttribute(eigh_backtransform_wy_kernel<16, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8 * 16 * 32 * 8));
      CUDA_TRY(cudaFuncSetAttribute(eigh_backtransform_wy_kernel<32, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8 * 32 * 32 * 8));
      CUDA_TRY(cudaFuncSetAttribute(eigh_backtransform_wy_kernel<64, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4 * 64 * 32 * 8));
      CUDA_TRY(cudaFuncSetAttribute(eigh_backtransform_dmma_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8 * (16 * 64 + 8) * 8));
      CUDA_TRY(cudaFuncSetAttribute(eigh_backtransform_dmma_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8 * (8 * 64 + 8) * 8));
      CUDA_TRY(cudaFuncSetAttribute(eigh_backtransform_dmma_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (32 * 64 + 8) * 8));
      attr_dev = g.device;
    }
    const int rows_per_cta = EIGH_SYTRD_THREADS / 32;
    int G = (n + rows_per_cta - 1) / rows_per_cta;
    void* args[] = {&P};
    const bool reg_path = G <= g.num_sms && G <= 160 && n <= 37 * 32 && !getenv("PLB200_EIGH_NO_REG");
    if (reg_path) {  // one row per warp, register-resident
      const void* fn = n <= 256   ? (const void*)eigh_sytrd_reg_kernel<8>
                       : n <= 512 ? (const void*)eigh_sytrd_reg_kernel<16>
                       : n <= 768 ? (const void*)eigh_sytrd_reg_kernel<24>
                       : n <= 1024 ? (const void*)eigh_sytrd_reg_kernel<32>
                                   : (const void*)eigh_sytrd_reg_kernel<37>;
      CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(EIGH_SYTRD_THREADS), args, 0, st));
    } else {  // rows in L2, 12 warps per CTA
      G = (n + EIGH_L2_WARPS - 1) / EIGH_L2_WARPS;
      if (G > g.num_sms) G = g.num_sms;
      const int rows_per_cta = (n + G - 1) / G;  // <= 14 at n <= 2048
      if (n <= 2048 && !getenv("PLB200_EIGH_NO_SMR"))
        CUDA_TRY(cudaLaunchCooperativeKernel((const void*)eigh_sytrd_l2_kernel<2, true>, dim3(G), dim3(EIGH_L2_THREADS), args,
                                             (size_t)(4 * 2048 + rows_per_cta * 1024) * 8, st));
      else if (n <= 2048)
        CUDA_TRY(cudaLaunchCooperativeKernel((const void*)eigh_sytrd_l2_kernel<2, false>, dim3(G), dim3(EIGH_L2_THREADS), args, (size_t)4 * 2048 * 8, st));
      else
        CUDA_TRY(cudaLaunchCooperativeKernel((const void*)eigh_sytrd_l2_kernel<4, false>, dim3(G), dim3(EIGH_L2_THREADS), args, (size_t)4 * 4096 * 8, st));
    }
    LAUNCH_CHECK();
  }
  if (getenv("PLB200_EIGH_PROF")) {
    static long long prof[8 * 160];
    CUDA_TRY(cudaMemcpyAsync(prof, d_rho, sizeof(prof), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int c : {0, 1, 73, 100})
      fprintf(stderr, "sytrd prof cta %d warp 6: loads %lld elementwise %lld ctabar %lld rowpass %lld publish %lld gridbar %lld (%lld) cycles\n", c, prof[8 * c],
              prof[8 * c + 1], prof[8 * c + 2], prof[8 * c + 3], prof[8 * c + 4], prof[8 * c + 5], prof[8 * c + 6]);
  }
  // 2. divide and conquer
  dc_leaf_kernel<<<(unsigned)T.leaves.size(), 256, 0, st>>>(d_d, d_e, n, (const int2*)(dtab + at_leaves), d_lam, QA, ldq, d_scal);
  LAUNCH_CHECK();
  for (auto& a : lat) {
    const EighMerge* dm = (const EighMerge*)(dtab + a.merges);
    DeflateParams D;
    D.merges = dm; D.e = d_e; D.scal = d_scal; D.lam = d_lam; D.Q = QA; D.ldq = ldq; D.perm = d_perm; D.rot = d_rot; D.nrot = d_nrot; D.Kp = d_Kp;
    D.rho = d_rho; D.dl = d_dl; D.wz = d_wz;
    const int dthreads = a.kmax >= 1024 ? 1024 : ((a.kmax + 31) & ~31);
    dc_deflate_kernel<<<a.nm, dthreads, (size_t)48 * a.kmax, st>>>(D);
    LAUNCH_CHECK();
    dc_permute_kernel<<<a.nrows, 256, (size_t)a.kmax * 8, st>>>(dm, dtab + a.r2m, dtab + a.rows, QA, QB, ldq, d_perm, d_rot, d_nrot);
    LAUNCH_CHECK();
    {
      const dim3 sg((a.kmax + 7) / 8, a.nm);
      if (a.kmax <= 64) dc_secular_kernel<2><<<sg, 256, 0, st>>>(dm, d_Kp, d_rho, d_dl, d_wz, d_lam, VT, ldq);
      else if (a.kmax <= 128) dc_secular_kernel<4><<<sg, 256, 0, st>>>(dm, d_Kp, d_rho, d_dl, d_wz, d_lam, VT, ldq);
      else if (a.kmax <= 256) dc_secular_kernel<8><<<sg, 256, 0, st>>>(dm, d_Kp, d_rho, d_dl, d_wz, d_lam, VT, ldq);
      else if (a.kmax <= 512) dc_secular_kernel<16><<<sg, 256, 0, st>>>(dm, d_Kp, d_rho, d_dl, d_wz, d_lam, VT, ldq);
      else if (a.kmax <= 1024) dc_secular_kernel<32><<<sg, 256, 0, st>>>(dm, d_Kp, d_rho, d_dl, d_wz, d_lam, VT, ldq);
      else dc_secular_kernel<0><<<sg, 256, 0, st>>>(dm, d_Kp, d_rho, d_dl, d_wz, d_lam, VT, ldq);
    }
    LAUNCH_CHECK();
    dc_lowner_kernel<<<dim3((a.kmax + 31) / 32, a.nm), 256, 0, st>>>(dm, d_Kp, d_dl, d_wz, VT, ldq, d_what);
    LAUNCH_CHECK();
    dc_vectors_kernel<<<dim3((a.kmax + 7) / 8, a.nm), 256, 0, st>>>(dm, d_Kp, d_what, VT, ldq);
    LAUNCH_CHECK();
    // Q block <- (permuted Q block) V : C[i][j] = sum_l QB[i][l] VT[j][l] on the DMMA Gram engine
    const std::vector<EighMerge>& lv = T.levels[&a - lat.data()];
    if (a.kmax <= EIGH_SMALL_GEMM_MAX) {  // every merge of the level in one launch
      const int tl = (a.kmax + 31) / 32;
      dc_small_gemm_kernel<<<dim3(tl * tl, a.nm), 256, (size_t)2 * 32 * (a.kmax | 1) * 8, st>>>(dm, QB, VT, QA, ldq);
      LAUNCH_CHECK();
      continue;
    }
    for (const EighMerge& mg : lv) {
      const int k = mg.hi - mg.lo;
      const size_t off = (size_t)mg.lo * ldq + mg.lo;
      PL_TRY(gram_dev_impl(PL_KERNEL_LINEAR, QB + off, k, ldq, VT + off, k, ldq, k, 1, 0, nullptr, 1.0, 1.0, QA + off, ldq, PL_FILL_FULL, 0, 1, 0, st));
    }
  }
  // 3. order, transpose into the caller's matrix, back-transformation
  eigh_rank_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_lam, n, d_scal, d_order, dW);
  LAUNCH_CHECK();
  if (want_vectors) {
    eigh_gather_transpose_kernel<<<dim3((n + 31) / 32, (n + 31) / 32), 256, 0, st>>>(QA, ldq, n, d_order, dA, lda);
    LAUNCH_CHECK();
    if (n <= 2048 && !getenv("PLB200_EIGH_NO_REG")) {
      const unsigned blocks = (unsigned)((n + 7) / 8);
      if (n >= 3 && !getenv("PLB200_EIGH_SCALAR_BT")) {  // DMMA form
        eigh_larft_kernel<8><<<(unsigned)((n - 2 + 7) / 8), 256, 0, st>>>(S, ldq, d_tau, n, d_Tf);
        LAUNCH_CHECK();
        if (n <= 256) eigh_backtransform_dmma_kernel<4, true><<<blocks, 256, (size_t)2 * 8 * (4 * 64 + 8) * 8, st>>>(S, ldq, d_Tf, n, dA, lda);
        else if (n <= 512) eigh_backtransform_dmma_kernel<8, true><<<blocks, 256, (size_t)2 * 8 * (8 * 64 + 8) * 8, st>>>(S, ldq, d_Tf, n, dA, lda);
        else if (n <= 1024) eigh_backtransform_dmma_kernel<16, true><<<blocks, 256, (size_t)2 * 8 * (16 * 64 + 8) * 8, st>>>(S, ldq, d_Tf, n, dA, lda);
        else eigh_backtransform_dmma_kernel<32, false><<<blocks, 256, (size_t)8 * (32 * 64 + 8) * 8, st>>>(S, ldq, d_Tf, n, dA, lda);
      } else if (n >= 3) {
#define PLB_BT(NCH, R, NV)                                                                                                          \
  do {                                                                                                                              \
    eigh_larft_kernel<R><<<(unsigned)((n - 2 + R - 1) / R), 256, 0, st>>>(S, ldq, d_tau, n, d_Tf);                                  \
    LAUNCH_CHECK();                                                                                                                 \
    eigh_backtransform_wy_kernel<NCH, R, NV><<<blocks, 256 / NV, (size_t)2 * R * NCH * 32 * 8, st>>>(S, ldq, d_Tf, n, dA, lda);     \
  } while (0)
        if (n <= 256) PLB_BT(8, 8, 1);
        else if (n <= 512) PLB_BT(16, 8, 1);
        else if (n <= 1024) PLB_BT(32, 8, 1);
        else PLB_BT(64, 4, 1);
#undef PLB_BT
      }
    } else {
      int nw = (int)((200 * 1024) / ((size_t)n * 8));
      nw = nw > 8 ? 8 : (nw < 1 ? 1 : nw);
      eigh_backtransform_kernel<<<(n + nw - 1) / nw, nw * 32, (size_t)nw * n * 8, st>>>(S, ldq, d_tau, n, dA, lda);
    }
    LAUNCH_CHECK();
  }
  unsigned long long flags[2] = {0, 0};
  CUDA_TRY(cudaMemcpyAsync(flags + 1, d_bar + EIGH_BAR_ABORT, 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (flags[1] != 0) return set_err(PL_ECUDA, "eigensolver: a grid barrier of the tridiagonalisation timed out");
  return PL_OK;
}

// In place: the `fill` triangle (row-major view) of dA (n x n, leading dimension lda) is read; on return row k of dA is the unit eigenvector of
// the k-th smallest eigenvalue dW[k] (column k in Julia's column-major view). Synchronises `st` (cuSOLVER reports through a device flag).
int eigh_dev_impl(double* dA, int64_t n, int64_t lda, int fill, int want_vectors, double* dW, cudaStream_t st) {
  PL_TRY(need_ready());
  if (n <= 0 || lda < n || !dA || !dW) return set_err(PL_EINVAL, "bad arguments");
  if (fill < 1 || fill > 3) return set_err(PL_EINVAL, "bad fill");
  {
    const char* mode = getenv("PLB200_EIGH");
    const char* mx = getenv("PLB200_EIGH_OWN_MAX");
    int64_t own_max = mx ? atoll(mx) : 2048;  // <= 1184: register-resident tridiagonalisation (2.7x cuSOLVER); <= 2048: rows in L2 (1.6x)
    if (own_max > EIGH_MAX_N) own_max = EIGH_MAX_N;
    const bool force_own = mode && !strcmp(mode, "own"), force_lib = mode && !strcmp(mode, "cusolver");
    if (force_own) return eigh_own_impl(dA, n, lda, fill, want_vectors, dW, st);
    if (!force_lib && n <= own_max) {
      const int rc = eigh_own_impl(dA, n, lda, fill, want_vectors, dW, st);
      // automatic routing only: a device that cannot run the cooperative kernels (launch refused, barrier timed out -- the input is
      // untouched until the last step) hands the matrix to cuSOLVER, and says so once; bad input is reported as it is
      if (rc == PL_OK || rc == PL_EINVAL) return rc;
      static bool told = false;
      if (!told) {
        fprintf(stderr, "libplb200: built-in eigensolver failed (%s); falling back to cuSOLVER for this process\n", pl_last_error());
        told = true;
      }
      cudaGetLastError();
    }
  }
  PL_TRY(cusolver_ready());
  cusolverStatus_t r = cs.SetStream(cs.h, st);
  if (r != CUSOLVER_STATUS_SUCCESS) return set_err(PL_ECUDA, "cusolverDnSetStream failed with status %d", (int)r);
  // row-major [i][j], j <= i (PL_FILL_RM_LOWER) is column-major element (row j, col i), j <= i: the UPPER triangle for cuSOLVER
  const cublasFillMode_t uplo = (fill & PL_FILL_RM_LOWER) ? CUBLAS_FILL_MODE_UPPER : CUBLAS_FILL_MODE_LOWER;
  const cusolverEigMode_t jobz = want_vectors ? CUSOLVER_EIG_MODE_VECTOR : CUSOLVER_EIG_MODE_NOVECTOR;
  size_t wdev = 0, whost = 0;
  r = cs.XsyevdBufferSize(cs.h, cs.params, jobz, uplo, n, CUDA_R_64F, dA, lda, CUDA_R_64F, dW, CUDA_R_64F, &wdev, &whost);
  if (r != CUSOLVER_STATUS_SUCCESS) return set_err(PL_ECUDA, "cusolverDnXsyevd_bufferSize failed with status %d (n = %lld)", (int)r, (long long)n);
  void *dwork, *dinfo;
  PL_TRY(ws_get(WS_EIG_WORK, wdev, &dwork));
  PL_TRY(ws_get(WS_EIG_INFO, 16, &dinfo));
  void* hwork = whost ? malloc(whost) : nullptr;
  if (whost && !hwork) return set_err(PL_ENOMEM, "host workspace of %zu bytes", whost);
  r = cs.Xsyevd(cs.h, cs.params, jobz, uplo, n, CUDA_R_64F, dA, lda, CUDA_R_64F, dW, CUDA_R_64F, dwork, wdev, hwork, whost, (int*)dinfo);
  int info = -1;
  cudaError_t e = cudaMemcpyAsync(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  free(hwork);
  if (r != CUSOLVER_STATUS_SUCCESS) return set_err(PL_ECUDA, "cusolverDnXsyevd failed with status %d", (int)r);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_err(PL_ECUDA, "eigensolver: %s", cudaGetErrorString(e));
  }
  if (info != 0) return set_err(PL_ECUDA, "cusolverDnXsyevd did not converge / bad argument (info = %d)", info);
  return PL_OK;
}

// the resident L-ensemble of the last pl_ell_ensemble call
struct EllState {
  int64_t n = 0, ld = 0;
  bool valid = false;
  bool has_kernel = false;  // a full symmetric copy of the kernel matrix is kept next to its eigenvectors (greedy MAP needs rows of K)
} ell_state;

int inclusion_prob_dev_impl(const double* dV, int64_t nvec, int64_t n, int64_t ldv, const double* dgamma, double* dp, cudaStream_t st) {
  int64_t rpb = nvec / 64;
  rpb = rpb < 16 ? 16 : (rpb + 1) / 2 * 2;
  const int64_t nblk = (nvec + rpb - 1) / rpb;
  void* part = nullptr;
  PL_TRY(ws_get(WS_COLSUM_PARTIAL, (size_t)nblk * n * 8, &part));
  dim3 grid((unsigned)((n + 127) / 128), (unsigned)nblk);
  weighted_colsumsq_partial_kernel<<<grid, 128, 0, st>>>(dV, nvec, n, ldv, dgamma, (int)rpb, (double*)part);
  LAUNCH_CHECK();
  colsum_final_kernel<<<(unsigned)((n + 31) / 32), 256, 0, st>>>((const double*)part, nblk, n, 0.0, dp);
  LAUNCH_CHECK();
  return PL_OK;
}


// resident accumulator of the out-of-core normal equations (pl_neq_begin / pl_neq_add / pl_neq_finish)
struct NeqState {
  int64_t Kf = 0, rows = 0, batches = 0;
  bool open = false;
} neq_state;


// One-pass centred second moment about a SHIFT z close to the mean: S_z = sum_r (x_r - z)(x_r - z)', rho = sum_r (x_r - z).
// With the true mean m = z + rho / n the centred sum is S_z - rho (m-z)' - (m-z) rho' + n (m-z)(m-z)'; because z is already within
// sigma / sqrt(prefix) of m the correction is ~1e-5 of S and nothing cancels — the mean pass over the data is not needed.
int cov_shifted_dev_impl(const double* dA, int64_t R, int64_t Kf, int64_t lda, const double* dz, double* dS, double* drho, cudaStream_t st) {
  if (!dz || !drho) return set_err(PL_EINVAL, "null pointer");
  return syrk_dev_impl(dA, R, Kf, lda, nullptr, 0, 1.0, 1.0, nullptr, dz, dS, nullptr, st, drho);
}

int cov_recentre_dev_impl(double* dS, int64_t Kf, const double* drho, const double* ddelta, double n_rows, cudaStream_t st) {
  PL_TRY(need_ready());
  if (!dS || !drho || !ddelta || Kf <= 0) return set_err(PL_EINVAL, "bad arguments");
  int blocks = (int)((Kf * Kf + 255) / 256);
  if (blocks > g.num_sms * 8) blocks = g.num_sms * 8;
  cov_recentre_kernel<<<blocks, 256, 0, st>>>(dS, Kf, drho, ddelta, n_rows);
  LAUNCH_CHECK();
  return PL_OK;
}


int metrics_dev_impl(const double* dxp, const double* dx, int64_t n, double* dout /*4*/, cudaStream_t st) {
  PL_TRY(need_ready());
  if (!dxp || !dx || !dout || n <= 0) return set_err(PL_EINVAL, "bad arguments");
  int nb = (int)((n + 255) / 256);
  if (nb > g.num_sms * 4) nb = g.num_sms * 4;
  void* ws;
  PL_TRY(ws_get(WS_MET, (size_t)(nb * 6 + 1) * 8, &ws));
  double* part = (double*)ws;
  double* mean = part + nb * 6;
  metrics_partial_kernel<<<nb, 256, 0, st>>>(dxp, dx, n, nullptr, part);
  LAUNCH_CHECK();
  metrics_final_kernel<<<1, 32, 0, st>>>(part, nb, n, 0, mean);
  LAUNCH_CHECK();
  metrics_partial_kernel<<<nb, 256, 0, st>>>(dxp, dx, n, mean, part);
  LAUNCH_CHECK();
  metrics_final_kernel<<<1, 32, 0, st>>>(part, nb, n, 1, dout);
  LAUNCH_CHECK();
  return PL_OK;
}

}  // namespace

namespace {

// defined in multi.inl (included at the end of this file): single-process multi-GPU forms of the host entry points
void multi_shutdown();
void gram_graphs_clear(int ctx_index);  // cached CUDA graphs of pl_gram_dev (defined next to it)
int normal_eq_multi(const double* const* blocks, const int64_t* rows, const int64_t* lds, int64_t nblocks, int64_t R, int64_t Kf, const double* w,
                    int64_t ne, double w_e, double w_f, const double* b, int ic, double lambda, double* AtWA, double* AtWb, int64_t chunk_rows);
int cov_multi(const double* X, int64_t n, int64_t Kf, int64_t ldx, double* mean_inout, int mean_given, int center, double* C);
int gram_sym_multi(int kernel_id, const double* X, int64_t n, int64_t d, int64_t ldx, int alpha, int cinv_kind, const double* cinv, double ell,
                   double beta, double* K, int64_t ldk);
// inputs below these sizes stay on device 0 even when the library drives several GPUs
constexpr double MULTI_MIN_ROW_BYTES = 256.0 * 1048576.0;  // normal equations / covariance: bytes of the row matrix
constexpr int64_t MULTI_MIN_GRAM_N = 8192;                 // symmetric kernel matrices: number of features

// Symmetric(K) uplo = :U storage (kernels.jl:217-222) streamed to the host: row block jb of the row-major view is complete as soon as
// tile column jb is. The T tile columns are cut into `nchunk_total` ranges of equal work; this call computes ranges [c_lo, c_hi) — all of
// them on one GPU, one contiguous share per device when several GPUs work on one matrix — launching the Gram kernel once per range and
// copying each finished range to the host on the copy stream while the next ranges compute: the PCIe copy (the e2e bottleneck)
// starts at once. Only rows [128 bounds[c_lo], 128 bounds[c_hi]) of the result exist on this device. Records g.ev[1..3].
int gram_sym_stream(int kernel_id, const double* d1, int64_t n1, int64_t l1, int64_t d, int alpha, int cinv_kind, const double* dc, double ell,
                    double beta, double* K, int64_t ldk, int c_lo, int c_hi, int nchunk_total) {
  const int T = (int)((n1 + GRAM_BM - 1) / GRAM_BM);
  const int64_t ldo = (n1 + 1) & ~(int64_t)1;
  std::vector<int> bounds(nchunk_total + 1);
  for (int c = 0; c <= nchunk_total; ++c) bounds[c] = (int)(T * sqrt((double)c / nchunk_total) + 0.5);
  if (c_lo == 0 && c_hi == nchunk_total && nchunk_total == 8 && !getenv("PLB200_STREAM_EQUAL")) {
    // One device does all ranges: the copy is ~8 x slower than the compute, so what matters is how early it STARTS. The first range is
    // one wave of tiles (its copy begins ~50 us after the first launch instead of after an eighth of the matrix), the next ones grow
    // geometrically and compute stays ahead of the copy throughout (PLB200_STREAM_TRACE=1 prints the timeline).
    static const double frac[9] = {0.0, 0.012, 0.06, 0.2, 0.4, 0.6, 0.8, 0.9, 1.0};
    for (int c = 0; c <= 8; ++c) bounds[c] = (int)(T * sqrt(frac[c]) + 0.5);
  }
  bounds[0] = 0;
  bounds[nchunk_total] = T;
  const int64_t row_lo = (int64_t)bounds[c_lo] * GRAM_BM;
  int64_t row_hi = (int64_t)bounds[c_hi] * GRAM_BM;
  if (row_hi > n1) row_hi = n1;
  void* outbuf;
  PL_TRY(ws_get(WS_H_OUT1, (size_t)(row_hi > row_lo ? row_hi - row_lo : 1) * ldo * 8, &outbuf));
  double* dK = (double*)outbuf - row_lo * ldo;  // rows below row_lo are never addressed
  CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
  GramLaunch L;
  PL_TRY(gram_dev_impl(kernel_id, d1, n1, l1, nullptr, n1, 0, d, alpha, cinv_kind, dc, ell, beta, dK, ldo, PL_FILL_RM_LOWER, 0, 1, 0, g.stream, &L));
  const int nch = c_hi - c_lo;
  if (nch > 12) return set_err(PL_EINVAL, "too many chunks per device");
  // PLB200_STREAM_TRACE=1: when each range's kernel and each range's copy finished, in ms after the first launch, on stderr
  const bool trace = getenv("PLB200_STREAM_TRACE") != nullptr;
  cudaEvent_t tk[12] = {}, tc[12] = {};
  if (trace)
    for (int c = 0; c < nch; ++c) {
      CUDA_TRY(cudaEventCreate(&tk[c]));
      CUDA_TRY(cudaEventCreate(&tc[c]));
    }
  for (int c = 0; c < nch; ++c) {
    if (bounds[c_lo + c + 1] > bounds[c_lo + c]) {
      L.tcol_begin = bounds[c_lo + c];
      L.tcol_end = bounds[c_lo + c + 1];
      PL_TRY(launch_gram(L, g.stream));
    }
    CUDA_TRY(cudaEventRecord(g.chunk_ev[c], g.stream));
    if (trace) CUDA_TRY(cudaEventRecord(tk[c], g.stream));
  }
  CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
  const bool stage_out = staging_for(K, (size_t)n1 * n1 * 4);
  for (int c = 0; c < nch; ++c) {
    CUDA_TRY(cudaStreamWaitEvent(g.copy_stream, g.chunk_ev[c], 0));
    for (int tb = bounds[c_lo + c]; tb < bounds[c_lo + c + 1]; ++tb) {
      const int64_t r0 = (int64_t)tb * GRAM_BM;
      const int64_t r1 = r0 + GRAM_BM < n1 ? r0 + GRAM_BM : n1;
      if (stage_out)
        CUDA_TRY(g_stager.d2h_2d(K + r0 * ldk, (size_t)ldk * 8, dK + r0 * ldo, (size_t)ldo * 8, (size_t)r1 * 8, (size_t)(r1 - r0), g.copy_stream, false));
      else
        CUDA_TRY(cudaMemcpy2DAsync(K + r0 * ldk, (size_t)ldk * 8, dK + r0 * ldo, (size_t)ldo * 8, (size_t)r1 * 8, (size_t)(r1 - r0), cudaMemcpyDeviceToHost, g.copy_stream));
    }
    if (trace) CUDA_TRY(cudaEventRecord(tc[c], g.copy_stream));
  }
  if (stage_out) CUDA_TRY(g_stager.finish());
  CUDA_TRY(cudaEventRecord(g.ev[3], g.copy_stream));
  CUDA_TRY(cudaStreamSynchronize(g.copy_stream));
  if (trace) {
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    fprintf(stderr, "[plb200 stream trace] range: kernel done / copy done (ms after the first launch):");
    for (int c = 0; c < nch; ++c) {
      float a = 0, b = 0;
      cudaEventElapsedTime(&a, g.ev[1], tk[c]);
      cudaEventElapsedTime(&b, g.ev[1], tc[c]);
      fprintf(stderr, " %d: %.2f / %.2f;", c, a, b);
      cudaEventDestroy(tk[c]);
      cudaEventDestroy(tc[c]);
    }
    fprintf(stderr, "\n");
  }
  return finish_timing();
}

// create context `idx` on CUDA device `device`: streams, events, kernel attributes. Leaves the calling thread on that device.
int ctx_create(int idx, int device) {
  Ctx& c = g_ctx[idx];
  if (c.ready) return PL_OK;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return set_err(PL_ENODEV, "device %d is sm_%d%d; libplb200 is built for sm_100a (B200) only", device, prop.major, prop.minor);
  CUDA_TRY(cudaSetDevice(device));
  c.device = device;
  c.index = idx;
  c.num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) return set_err(PL_ECUDA, "cuTensorMapEncodeTiled entry point not found");
  c.encode = (PFN_cuTensorMapEncodeTiled)fn;
  CUDA_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
  for (auto& ev : c.ev) CUDA_TRY(cudaEventCreate(&ev));
  for (auto& ev : c.chunk_ev) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
#define GRAM_ATTR(EPI_)                                                                                                                  \
  CUDA_TRY(cudaFuncSetAttribute(gram_kernel<EPI_, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM_BYTES + (EPI_ == EPI_RBF ? EXP_TAB_BYTES : 0))); \
  CUDA_TRY(cudaFuncSetAttribute(gram_kernel<EPI_, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM_BYTES + (EPI_ == EPI_RBF ? EXP_TAB_BYTES : 0))); \
  CUDA_TRY(cudaFuncSetAttribute(gram_kernel<EPI_, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM_BYTES + (EPI_ == EPI_RBF ? EXP_TAB_BYTES : 0))); \
  CUDA_TRY(cudaFuncSetAttribute(gram_kernel<EPI_, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM_BYTES_TMAST + (EPI_ == EPI_RBF ? EXP_TAB_BYTES : 0)));
  GRAM_ATTR(EPI_LINEAR)
  GRAM_ATTR(EPI_DOT)
  GRAM_ATTR(EPI_RBF)
  GRAM_ATTR(EPI_IMQ)
#undef GRAM_ATTR
  CUDA_TRY(cudaFuncSetAttribute(syrk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SYRK_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(syrk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SYRK_SMEM_BYTES));
  {
    // correctly rounded 2^(j/256): long double exp2l carries 64 bits, the rounding to double is the only error (<= 0.5 ulp)
    std::vector<double> tab(EXP_TAB_N);
    for (int j = 0; j < EXP_TAB_N; ++j) tab[j] = (double)exp2l((long double)j / (long double)EXP_TAB_N);
    CUDA_TRY(cudaMalloc(&c.exp_tab, EXP_TAB_N * 8));
    CUDA_TRY(cudaMemcpy(c.exp_tab, tab.data(), EXP_TAB_N * 8, cudaMemcpyHostToDevice));
  }
  c.ready = true;
  return PL_OK;
}

void ctx_destroy(int idx) {
  Ctx& c = g_ctx[idx];
  if (!c.ready) return;
  cudaSetDevice(c.device);
  cudaDeviceSynchronize();
  for (int i = 0; i < WS_NSLOTS; ++i) {
    if (c.ws[i]) cudaFree(c.ws[i]);
    c.ws[i] = nullptr;
    c.ws_bytes[i] = 0;
  }
  for (auto& ev : c.ev) {
    if (ev) cudaEventDestroy(ev);
    ev = nullptr;
  }
  for (auto& ev : c.chunk_ev) {
    if (ev) cudaEventDestroy(ev);
    ev = nullptr;
  }
  if (c.exp_tab) cudaFree(c.exp_tab);
  c.exp_tab = nullptr;
  c.stager.shutdown();
  gram_graphs_clear(idx);
  if (c.sched.ready) cudaEventDestroy(c.sched.ready);
  c.sched.ready = nullptr;
  c.sched.dev = nullptr;
  c.sched.R = -1;
  if (c.stream) cudaStreamDestroy(c.stream);
  if (c.copy_stream) cudaStreamDestroy(c.copy_stream);
  c.stream = nullptr;
  c.copy_stream = nullptr;
  c.ready = false;
}

int device_count_checked(int* count) {
  *count = 0;
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess || *count == 0) {
    cudaGetLastError();
    return set_err(PL_ENODEV, "no CUDA device available (%s); plb200 has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  return PL_OK;
}

}  // namespace

extern "C" {

int pl_init(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_ctx[0].ready) return PL_OK;
  int count = 0;
  PL_TRY(device_count_checked(&count));
  if (device < 0) CUDA_TRY(cudaGetDevice(&device));
  if (device >= count) return set_err(PL_EINVAL, "device %d out of range (%d devices)", device, count);
  tl_ctx = &g_ctx[0];
  PL_TRY(ctx_create(0, device));
  g_ndev = 1;
  return PL_OK;
}

void pl_finalize(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_ctx[0].ready) return;
  multi_shutdown();
  if (cs.h) {
    cudaSetDevice(g_ctx[0].device);
    if (cs.params) cs.DestroyParams(cs.params);
    cs.Destroy(cs.h);
    cs.h = nullptr;
    cs.params = nullptr;
  }
  ell_state.valid = false;
  neq_state.open = false;
  for (int i = PL_MAX_DEV - 1; i >= 0; --i) ctx_destroy(i);
  g_ndev = 0;
  tl_ctx = &g_ctx[0];
}

const char* pl_last_error(void) { return g_err; }

int pl_last_timing(double* kernel_ms, double* h2d_ms, double* d2h_ms) {
  if (kernel_ms) *kernel_ms = g.t_kernel;
  if (h2d_ms) *h2d_ms = g.t_h2d;
  if (d2h_ms) *d2h_ms = g.t_d2h;
  return PL_OK;
}

int pl_last_host_ms(double* ms) {
  if (ms) *ms = g.t_host;
  return PL_OK;
}

int64_t pl_launch_count(void) { return g_launches.load(); }

int pl_set_profiling(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  g.profiling = on != 0;
  g.main_timed = false;
  return PL_OK;
}

int pl_last_main_kernel_ms(double* ms) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!g.main_timed) return set_err(PL_EINVAL, "no dominant kernel has been timed: call pl_set_profiling(1) first");
  CUDA_TRY(cudaEventSynchronize(g.ev[5]));
  float t = 0;
  CUDA_TRY(cudaEventElapsedTime(&t, g.ev[4], g.ev[5]));
  if (ms) *ms = t;
  return PL_OK;
}

int64_t pl_workspace_bytes(void) {
  int64_t tot = 0;
  for (int c = 0; c < PL_MAX_DEV; ++c)
    for (int i = 0; i < WS_NSLOTS; ++i) tot += (int64_t)g_ctx[c].ws_bytes[i];
  return tot;
}

// Hands the library's device scratch back to the driver (it only ever grows otherwise: a k = 10 000 draw at n = 20 000 leaves 6.4 GB of
// factors behind). keep_resident != 0 keeps what later calls depend on: the resident L-ensemble (eigenvectors, eigenvalues, kernel
// matrix) and an open normal-equation accumulator; 0 drops those too (pl_ell_* then ask for a new pl_ell_ensemble, pl_neq_add for a
// new pl_neq_begin).
int pl_workspace_trim(int keep_resident) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  CUDA_TRY(cudaSetDevice(g.device));
  CUDA_TRY(cudaDeviceSynchronize());
  for (int i = 0; i < WS_NSLOTS; ++i) {
    if (keep_resident) {
      if (ell_state.valid && (i == WS_EIG_A || i == WS_EIG_W || i == WS_ELL_K)) continue;
      if (neq_state.open && (i == WS_ACC_S || i == WS_ACC_V)) continue;
    }
    if (g.ws[i]) CUDA_TRY(cudaFree(g.ws[i]));
    g.ws[i] = nullptr;
    g.ws_bytes[i] = 0;
  }
  ++g.ws_gen;
  g.sched.dev = nullptr;
  for (int c = 1; c < g_ndev; ++c) {  // the other devices of a pl_init_all session hold scratch only
    Ctx& x = g_ctx[c];
    CUDA_TRY(cudaSetDevice(x.device));
    CUDA_TRY(cudaDeviceSynchronize());
    for (int i = 0; i < WS_NSLOTS; ++i) {
      if (x.ws[i]) CUDA_TRY(cudaFree(x.ws[i]));
      x.ws[i] = nullptr;
      x.ws_bytes[i] = 0;
    }
    x.sched.dev = nullptr;
  }
  CUDA_TRY(cudaSetDevice(g.device));
  if (!keep_resident) {
    ell_state.valid = false;
    neq_state.open = false;
  }
  return PL_OK;
}

}  // extern "C"

namespace {
// pl_gram_dev is launch bound on small problems (config 1: four launches and three tensor-map encodes for 6 us of GPU work; a config-2
// shard on 8 GPUs: 48 us of preludes and launch gaps around a 540 us kernel). The launches of a call are therefore captured into a CUDA
// graph the SECOND time the identical call is made (same pointers, shapes, parameters, workspace generation) and replayed from then on:
// one cudaGraphLaunch instead of the launch sequence. The kernels read whatever the buffers hold at replay time, exactly like the
// plain launches. PLB200_NO_GRAPH=1 disables it; profiling (pl_set_profiling) bypasses it.
struct GramGraphKey {
  int kernel_id, alpha, cinv_kind, fill, part, nparts, packed;
  const void *dX1, *dX2, *dcinv, *dOut;
  int64_t n1, ldx1, n2, ldx2, d, ldo;
  double ell, beta;
  uint64_t ws_gen;
  unsigned env;
  bool operator==(const GramGraphKey& o) const { return memcmp(this, &o, sizeof(*this)) == 0; }
};
struct GramGraphEntry {
  GramGraphKey key;
  cudaGraphExec_t exec = nullptr;
  int64_t launches = 0;
  uint64_t last_use = 0;
  bool uncacheable = false;
};
constexpr int GRAM_GRAPH_SLOTS = 8;
GramGraphEntry g_gram_graphs[PL_MAX_DEV][GRAM_GRAPH_SLOTS];
uint64_t g_gram_graph_clock = 0;

unsigned env_bits() {
  unsigned h = 0;
  const char* names[] = {"PLB200_GRAM_SMALL", "PLB200_NO_PDL", "PLB200_GRAM_TMAST", "PLB200_GRAM_NOSTORE"};
  for (const char* nm : names) {
    const char* e = getenv(nm);
    h = h * 31u + (e ? (unsigned)(unsigned char)e[0] + 1u : 0u);
  }
  return h;
}

void gram_graphs_clear(int ctx_index) {
  for (auto& e : g_gram_graphs[ctx_index]) {
    if (e.exec) cudaGraphExecDestroy(e.exec);
    e = GramGraphEntry();
  }
}
}  // namespace

extern "C" {

int pl_gram_dev(int kernel_id, const double* dX1, int64_t n1, int64_t ldx1, const double* dX2, int64_t n2, int64_t ldx2, int64_t d, int alpha,
                int cinv_kind, const double* dcinv, double ell, double beta, double* dOut, int64_t ldo, int fill, int part, int nparts, int packed,
                void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  cudaStream_t st = (cudaStream_t)stream;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  const bool plain = !g.ready || g.profiling || getenv("PLB200_NO_GRAPH") || cudaStreamIsCapturing(st, &cap) != cudaSuccess ||
                     cap != cudaStreamCaptureStatusNone;  // (a caller that is capturing itself gets the plain launches in ITS graph)
  if (plain) {
    cudaGetLastError();
    return gram_dev_impl(kernel_id, dX1, n1, ldx1, dX2, n2, ldx2, d, alpha, cinv_kind, dcinv, ell, beta, dOut, ldo, fill, part, nparts, packed, st);
  }
  GramGraphKey key;
  memset(&key, 0, sizeof(key));
  key.kernel_id = kernel_id; key.alpha = alpha; key.cinv_kind = cinv_kind; key.fill = fill; key.part = part; key.nparts = nparts; key.packed = packed;
  key.dX1 = dX1; key.dX2 = dX2; key.dcinv = dcinv; key.dOut = dOut;
  key.n1 = n1; key.ldx1 = ldx1; key.n2 = n2; key.ldx2 = ldx2; key.d = d; key.ldo = ldo;
  key.ell = ell; key.beta = beta;
  key.ws_gen = g.ws_gen;
  key.env = env_bits();
  GramGraphEntry* tab = g_gram_graphs[g.index];
  GramGraphEntry* hit = nullptr;
  GramGraphEntry* victim = &tab[0];
  for (int i = 0; i < GRAM_GRAPH_SLOTS; ++i) {
    if (tab[i].last_use && tab[i].key == key) hit = &tab[i];
    if (tab[i].last_use < victim->last_use) victim = &tab[i];
  }
  if (hit && hit->exec) {  // third and later identical calls: replay
    hit->last_use = ++g_gram_graph_clock;
    PL_TRY(need_ready());
    CUDA_TRY(cudaGraphLaunch(hit->exec, st));
    g_launches.fetch_add(hit->launches);
    return PL_OK;
  }
  if (!hit || hit->uncacheable) {  // first sight: plain launches (this is also where the workspace grows), remember the call
    const int rc = gram_dev_impl(kernel_id, dX1, n1, ldx1, dX2, n2, ldx2, d, alpha, cinv_kind, dcinv, ell, beta, dOut, ldo, fill, part, nparts, packed, st);
    if (!hit && rc == PL_OK) {
      if (victim->exec) cudaGraphExecDestroy(victim->exec);
      *victim = GramGraphEntry();
      key.ws_gen = g.ws_gen;  // (the call itself may have grown the workspace)
      victim->key = key;
      victim->last_use = ++g_gram_graph_clock;
    }
    return rc;
  }
  // second identical call: capture the launches on the library's stream, instantiate, and run the graph on the caller's stream
  hit->last_use = ++g_gram_graph_clock;
  PL_TRY(need_ready());
  const int64_t l0 = g_launches.load();
  cudaGraph_t graph = nullptr;
  int rc = PL_OK;
  if (cudaStreamBeginCapture(g.stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
    rc = gram_dev_impl(kernel_id, dX1, n1, ldx1, dX2, n2, ldx2, d, alpha, cinv_kind, dcinv, ell, beta, dOut, ldo, fill, part, nparts, packed, g.stream);
    const cudaError_t ee = cudaStreamEndCapture(g.stream, &graph);
    if (rc == PL_OK && ee == cudaSuccess && graph && g.ws_gen == key.ws_gen && cudaGraphInstantiate(&hit->exec, graph, 0) == cudaSuccess) {
      hit->launches = g_launches.load() - l0;
    } else {
      hit->exec = nullptr;
      hit->uncacheable = true;
    }
    if (graph) cudaGraphDestroy(graph);
  } else {
    hit->uncacheable = true;
  }
  cudaGetLastError();
  if (hit->exec) {
    CUDA_TRY(cudaGraphLaunch(hit->exec, st));
    return PL_OK;
  }
  g_launches.store(l0);
  return gram_dev_impl(kernel_id, dX1, n1, ldx1, dX2, n2, ldx2, d, alpha, cinv_kind, dcinv, ell, beta, dOut, ldo, fill, part, nparts, packed, st);
}

int64_t pl_gram_tile_count(int64_t n1, int64_t n2, int part, int nparts) {
  if (nparts < 1 || part < 0 || part >= nparts || n1 <= 0) return 0;
  const int sym = (n2 <= 0);
  const int T1 = (int)((n1 + GRAM_BM - 1) / GRAM_BM);
  const int T2 = sym ? T1 : (int)((n2 + GRAM_BN - 1) / GRAM_BN);
  const int64_t slots = gram_num_slots(T1, T2, sym);
  const int64_t mine = (slots - part + nparts - 1) / nparts;
  return mine > 0 ? mine : 0;
}

int pl_gram_tile_coords(int64_t n1, int64_t n2, int part, int nparts, int64_t k, int64_t* ti, int64_t* tj) {
  const int64_t cnt = pl_gram_tile_count(n1, n2, part, nparts);
  if (k < 0 || k >= cnt) return set_err(PL_EINVAL, "tile index %lld out of range (%lld)", (long long)k, (long long)cnt);
  const int sym = (n2 <= 0);
  const int T1 = (int)((n1 + GRAM_BM - 1) / GRAM_BM);
  const int T2 = sym ? T1 : (int)((n2 + GRAM_BN - 1) / GRAM_BN);
  int a = -1, b = -1;
  // slots of the folded enumeration that hold no tile (odd T only) report (-1, -1)
  if (!gram_slot_to_tile(T1, T2, sym, k * nparts + part, a, b)) a = b = -1;
  if (ti) *ti = a;
  if (tj) *tj = b;
  return PL_OK;
}

// The SYRK schedule as a table (host arithmetic only, no device needed): which CTA runs which (unit, stage range) piece and where its
// partial goes. mode 0 = waves, 1 = streamk; waves <= 0 / diag_cost <= 0 take the library defaults. Rows of `table`: (cta, unit,
// first_stage, nstages, slot). Used by the CPU tests of the scheduling invariants and by tools.
int pl_syrk_schedule(int64_t R, int64_t Kf, int num_sms, int mode, int waves, double diag_cost, int64_t capacity, int* table, int64_t* npieces,
                     int64_t* nctas) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (R < 0 || Kf <= 0 || num_sms <= 0 || (mode != 0 && mode != 1)) return set_err(PL_EINVAL, "bad arguments");
  SyrkSchedConfig cfg = syrk_sched_config();
  cfg.mode = mode;
  cfg.waves = waves > 0 ? waves : 20;
  if (diag_cost > 0.0) cfg.diag_cost = diag_cost;
  const int T = (int)((Kf + SYRK_BT - 1) / SYRK_BT);
  SyrkSchedule sc;
  build_syrk_schedule(R, T, syrk_num_units(T), num_sms, cfg, &sc);
  if (npieces) *npieces = (int64_t)sc.pieces.size();
  if (nctas) *nctas = sc.nctas;
  if (table) {
    if (capacity < (int64_t)sc.pieces.size()) return set_err(PL_EINVAL, "table holds %lld rows, the schedule has %zu", (long long)capacity, sc.pieces.size());
    for (int q = 0; q < sc.nctas; ++q)
      for (int pc = sc.cta_first[q]; pc < sc.cta_first[q + 1]; ++pc) {
        const SyrkPiece& x = sc.pieces[pc];
        int* row = table + (size_t)pc * 5;
        row[0] = q; row[1] = x.unit; row[2] = x.first_stage; row[3] = x.nstages; row[4] = x.slot;
      }
  }
  return PL_OK;
}

int pl_syrk_dev(const double* dA, int64_t R, int64_t Kf, int64_t lda, const double* dw, int64_t n_energy_rows, double w_e, double w_f,
                const double* db, const double* dmean, double* dS, double* dvec, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  return syrk_dev_impl(dA, R, Kf, lda, dw, n_energy_rows, w_e, w_f, db, dmean, dS, dvec, (cudaStream_t)stream);
}

int pl_normal_eq_finish_dev(const double* dS, const double* dvec, int64_t Kf, int add_intercept, double lambda, double* dAtWA, double* dAtWb,
                            void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  return finish_dev_impl(dS, dvec, Kf, add_intercept, lambda, dAtWA, dAtWb, (cudaStream_t)stream);
}

int pl_colsum_dev(const double* dA, int64_t R, int64_t Kf, int64_t lda, double* dsum, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (R < 0 || Kf <= 0 || lda < Kf || !dA || !dsum) return set_err(PL_EINVAL, "bad arguments");
  const double* A;
  int64_t ld;
  PL_TRY(tma_operand(dA, R, Kf, lda, WS_PAD1, (cudaStream_t)stream, &A, &ld));
  return colsum(A, R, Kf, ld, 0.0, dsum, (cudaStream_t)stream);
}

int pl_scale_div_dev(const double* dIn, double denom, int64_t count, double* dOut, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (count < 0 || !dIn || !dOut) return set_err(PL_EINVAL, "bad arguments");
  if (count == 0) return PL_OK;
  int blocks = (int)((count + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  scale_div_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dIn, denom, count, dOut);
  LAUNCH_CHECK();
  return PL_OK;
}

static int predict_dev_impl(const double* dA, int64_t R, int64_t Kf, int64_t lda, const double* dbeta, double beta0, int64_t ne, double* dy,
                            cudaStream_t st) {
  PL_TRY(need_ready());
  if (R < 0 || Kf <= 0 || lda < Kf || !dA || !dbeta || !dy) return set_err(PL_EINVAL, "bad arguments");
  if (R == 0) return PL_OK;
  const double* A;
  int64_t ld;
  PL_TRY(tma_operand(dA, R, Kf, lda, WS_PAD1, st, &A, &ld));
  const double* b;
  int64_t ldb;
  PL_TRY(tma_operand(dbeta, 1, Kf, (Kf + 1) & ~(int64_t)1, WS_PADC, st, &b, &ldb));
  int64_t blocks = (R * 32 + 255) / 256;
  const int64_t cap = (int64_t)g.num_sms * 16;
  if (blocks > cap) blocks = cap;
  predict_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(A, R, Kf, ld, b, beta0, ne, dy);
  LAUNCH_CHECK();
  return PL_OK;
}

int pl_predict_dev(const double* dA, int64_t R, int64_t Kf, int64_t lda, const double* dbeta, double beta0, int64_t n_energy_rows, double* dy,
                   void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  return predict_dev_impl(dA, R, Kf, lda, dbeta, beta0, n_energy_rows, dy, (cudaStream_t)stream);
}

int pl_predict(const double* A, int64_t R, int64_t Kf, int64_t lda, const double* beta, double beta0, int64_t n_energy_rows, double* y) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!A || !beta || !y || R < 0 || Kf <= 0 || lda < Kf) return set_err(PL_EINVAL, "bad arguments");
  if (R == 0) return PL_OK;
  CUDA_TRY(cudaSetDevice(g.device));
  CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
  double *dA, *db;
  int64_t ld;
  PL_TRY(h2d_matrix(A, R, Kf, lda, WS_H_IN1, &dA, &ld));
  PL_TRY(h2d_vector(beta, Kf, WS_H_IN2, &db));
  void* dy;
  PL_TRY(ws_get(WS_H_OUT2, (size_t)R * 8, &dy));
  CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
  PL_TRY(predict_dev_impl(dA, R, Kf, ld, db, beta0, n_energy_rows, (double*)dy, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
  CUDA_TRY(cudaMemcpyAsync(y, dy, (size_t)R * 8, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[3], g.stream));
  return finish_timing();
}

int pl_sym_scale_div_dev(const double* dIn, int64_t K, double denom, double* dOut, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (K <= 0 || !dIn || !dOut) return set_err(PL_EINVAL, "bad arguments");
  int blocks = (int)((K * K + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  sym_scale_div_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dIn, K, denom, dOut);
  LAUNCH_CHECK();
  return PL_OK;
}

// ------------------------------------------------ host-pointer entry points ------------------------------------------------

static int gram_host(int kernel_id, const double* X1, int64_t n1, int64_t ldx1, const double* X2, int64_t n2, int64_t ldx2, int64_t d, int alpha,
                     int cinv_kind, const double* cinv, double ell, double beta, double* K, int64_t ldk, int fill) {
  PL_TRY(need_ready());
  const bool sym = (X2 == nullptr);
  if (!X1 || !K || n1 < 0 || d <= 0 || ldx1 < d || (!sym && (n2 < 0 || ldx2 < d))) return set_err(PL_EINVAL, "bad arguments");
  if (sym) n2 = n1;
  if (ldk < n2) return set_err(PL_EINVAL, "ldk smaller than the number of columns");
  if (sym && (fill < 1 || fill > 3)) return set_err(PL_EINVAL, "bad fill");
  if (n1 == 0 || n2 == 0) return PL_OK;
  if (sym && fill == PL_FILL_RM_LOWER && g_ndev > 1 && n1 >= MULTI_MIN_GRAM_N)  // several GPUs: tile-column ranges per device, parallel D2H
    return gram_sym_multi(kernel_id, X1, n1, d, ldx1, alpha, cinv_kind, cinv, ell, beta, K, ldk);
  CUDA_TRY(cudaSetDevice(g.device));
  CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
  double *d1, *d2 = nullptr, *dc = nullptr, *dK;
  int64_t l1, l2 = 0;
  PL_TRY(h2d_matrix(X1, n1, d, ldx1, WS_H_IN1, &d1, &l1));
  if (!sym) PL_TRY(h2d_matrix(X2, n2, d, ldx2, WS_H_IN2, &d2, &l2));
  const bool dist_kernel = (kernel_id == PL_KERNEL_RBF || kernel_id == PL_KERNEL_IMQ);
  if (dist_kernel && cinv_kind == PL_CINV_DIAGONAL) {
    if (!cinv) return set_err(PL_EINVAL, "cinv is NULL");
    PL_TRY(h2d_vector(cinv, d, WS_H_IN3, &dc));
  } else if (dist_kernel && cinv_kind == PL_CINV_FULL) {
    if (!cinv) return set_err(PL_EINVAL, "cinv is NULL");
    PL_TRY(h2d_vector(cinv, d * d, WS_H_IN3, &dc));  // dense d x d; gram_dev_impl pads it if d is odd
  }
  const int T = (int)((n1 + GRAM_BM - 1) / GRAM_BM);
  if (sym && fill == PL_FILL_RM_LOWER && T >= 16)  // the reference's storage contract at scale: computed and copied out in overlapped column ranges
    return gram_sym_stream(kernel_id, d1, n1, l1, d, alpha, cinv_kind, dc, ell, beta, K, ldk, 0, 8, 8);
  const int64_t ldo = (n2 + 1) & ~(int64_t)1;
  void* outbuf;
  PL_TRY(ws_get(WS_H_OUT1, (size_t)n1 * ldo * 8, &outbuf));
  dK = (double*)outbuf;
  CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
  PL_TRY(gram_dev_impl(kernel_id, d1, n1, l1, d2, n2, l2, d, alpha, cinv_kind, dc, ell, beta, dK, ldo, fill, 0, 1, 0, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
  if (!sym || fill == PL_FILL_FULL) {
    if (staging_for(K, (size_t)n1 * n2 * 8))
      CUDA_TRY(g_stager.d2h_2d(K, (size_t)ldk * 8, dK, (size_t)ldo * 8, (size_t)n2 * 8, (size_t)n1, g.stream));
    else
      CUDA_TRY(cudaMemcpy2DAsync(K, (size_t)ldk * 8, dK, (size_t)ldo * 8, (size_t)n2 * 8, (size_t)n1, cudaMemcpyDeviceToHost, g.stream));
  } else {
    // one triangle only: copy the staircase of tile-row blocks (half the PCIe traffic). Inside diagonal tiles the
    // unused half was written as zeros by the kernel, matching the reference's zeros(n,n) storage.
    const int64_t edge = gram_small_tiles(n1, n2, true, 1, kernel_id) ? 64 : GRAM_BM;
    for (int64_t r0 = 0; r0 < n1; r0 += edge) {
      const int64_t r1 = r0 + edge < n1 ? r0 + edge : n1;
      if (fill == PL_FILL_RM_LOWER) {
        CUDA_TRY(cudaMemcpy2DAsync(K + r0 * ldk, (size_t)ldk * 8, dK + r0 * ldo, (size_t)ldo * 8, (size_t)r1 * 8, (size_t)(r1 - r0), cudaMemcpyDeviceToHost, g.stream));
      } else {
        CUDA_TRY(cudaMemcpy2DAsync(K + r0 * ldk + r0, (size_t)ldk * 8, dK + r0 * ldo + r0, (size_t)ldo * 8, (size_t)(n2 - r0) * 8, (size_t)(r1 - r0), cudaMemcpyDeviceToHost, g.stream));
      }
    }
  }
  CUDA_TRY(cudaEventRecord(g.ev[3], g.stream));
  return finish_timing();
}

int pl_gram_dot(const double* X, int64_t n, int64_t d, int64_t ldx, int alpha, double* K, int64_t ldk, int fill) {
  std::lock_guard<std::mutex> lk(g_mu);
  return gram_host(PL_KERNEL_DOT, X, n, ldx, nullptr, 0, 0, d, alpha, 0, nullptr, 1.0, 1.0, K, ldk, fill);
}

int pl_gram_rbf(const double* X, int64_t n, int64_t d, int64_t ldx, int cinv_kind, const double* cinv, double ell, double beta, double* K,
                int64_t ldk, int fill) {
  std::lock_guard<std::mutex> lk(g_mu);
  return gram_host(PL_KERNEL_RBF, X, n, ldx, nullptr, 0, 0, d, 0, cinv_kind, cinv, ell, beta, K, ldk, fill);
}

int pl_gram_imq(const double* X, int64_t n, int64_t d, int64_t ldx, int cinv_kind, const double* cinv, double c2, double ell, double* K,
                int64_t ldk, int fill) {
  std::lock_guard<std::mutex> lk(g_mu);
  return gram_host(PL_KERNEL_IMQ, X, n, ldx, nullptr, 0, 0, d, 0, cinv_kind, cinv, ell, c2, K, ldk, fill);
}

int pl_gram_cross(int kernel_id, const double* X1, int64_t n1, int64_t ldx1, const double* X2, int64_t n2, int64_t ldx2, int64_t d, int alpha,
                  int cinv_kind, const double* cinv, double ell, double beta, double* K, int64_t ldk) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!X2) return set_err(PL_EINVAL, "X2 is NULL (use pl_gram_dot / pl_gram_rbf for the symmetric form)");
  return gram_host(kernel_id, X1, n1, ldx1, X2, n2, ldx2, d, alpha, cinv_kind, cinv, ell, beta, K, ldk, PL_FILL_FULL);
}

int pl_gram_part(int kernel_id, const double* X, int64_t n, int64_t d, int64_t ldx, int alpha, int cinv_kind, const double* cinv, double ell,
                 double beta, int part, int nparts, double* tiles) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!X || !tiles || n < 0 || d <= 0 || ldx < d || nparts < 1 || part < 0 || part >= nparts) return set_err(PL_EINVAL, "bad arguments");
  if (n == 0) return PL_OK;
  CUDA_TRY(cudaSetDevice(g.device));
  CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
  double *dX, *dc = nullptr;
  int64_t ld;
  PL_TRY(h2d_matrix(X, n, d, ldx, WS_H_IN1, &dX, &ld));
  if ((kernel_id == PL_KERNEL_RBF || kernel_id == PL_KERNEL_IMQ) && cinv_kind != PL_CINV_IDENTITY) {
    if (!cinv) return set_err(PL_EINVAL, "cinv is NULL");
    PL_TRY(h2d_vector(cinv, cinv_kind == PL_CINV_DIAGONAL ? d : d * d, WS_H_IN3, &dc));
  }
  const int T = (int)((n + GRAM_BM - 1) / GRAM_BM);
  const int64_t slots = gram_num_slots(T, T, 1);
  const int64_t mine = (slots - part + nparts - 1) / nparts;
  void* outbuf;
  PL_TRY(ws_get(WS_H_OUT1, (size_t)(mine > 0 ? mine : 1) * GRAM_BM * GRAM_BN * 8, &outbuf));
  CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
  PL_TRY(gram_dev_impl(kernel_id, dX, n, ld, nullptr, 0, 0, d, alpha, cinv_kind, dc, ell, beta, (double*)outbuf, GRAM_BN, PL_FILL_FULL, part, nparts, 1, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
  if (mine > 0) CUDA_TRY(cudaMemcpyAsync(tiles, outbuf, (size_t)mine * GRAM_BM * GRAM_BN * 8, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[3], g.stream));
  return finish_timing();
}

int pl_normal_eq(const double* A, int64_t R, int64_t Kf, int64_t lda, const double* w, int64_t n_energy_rows, double w_e, double w_f,
                 const double* b, int add_intercept, double lambda, double* AtWA, double* AtWb) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!A || !AtWA || R < 0 || Kf <= 0 || lda < Kf) return set_err(PL_EINVAL, "bad arguments");
  if (AtWb && !b) return set_err(PL_EINVAL, "AtWb requested without b");
  CUDA_TRY(cudaSetDevice(g.device));
  if (g_ndev > 1 && (double)R * (double)Kf * 8.0 >= MULTI_MIN_ROW_BYTES) {
    // several GPUs (pl_init_all): rows sharded over the devices, every device streams its own shard over its own PCIe link, one all-reduce
    const double* blk[1] = {A};
    const int64_t rows[1] = {R}, lds[1] = {lda};
    const int64_t ne = n_energy_rows < 0 ? 0 : (n_energy_rows > R ? R : n_energy_rows);
    return normal_eq_multi(blk, rows, lds, 1, R, Kf, w, ne, w_e, w_f, b, add_intercept ? 1 : 0, lambda, AtWA, AtWb, 0);
  }
  if ((double)R * (double)Kf * 8.0 > 768.0 * 1048576.0) {
    // large inputs: stream row chunks through two device buffers, H2D overlapped with the SYRK (same code as pl_normal_eq_blocks)
    int64_t chunk_rows = (((int64_t)512 << 20) / (Kf * 8) + 15) / 16 * 16;
    const double* blk[1] = {A};
    const int64_t rows[1] = {R}, lds[1] = {lda};
    const int64_t ne = n_energy_rows < 0 ? 0 : (n_energy_rows > R ? R : n_energy_rows);
    return normal_eq_stream(blk, rows, lds, 1, R, Kf, w, ne, w_e, w_f, b, add_intercept ? 1 : 0, lambda, AtWA, AtWb, chunk_rows);
  }
  CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
  double *dA, *dw = nullptr, *db = nullptr;
  int64_t ld;
  PL_TRY(h2d_matrix(A, R, Kf, lda, WS_H_IN1, &dA, &ld));
  if (w) PL_TRY(h2d_vector(w, R, WS_H_IN2, &dw));
  if (b) PL_TRY(h2d_vector(b, R, WS_H_IN3, &db));
  const int ic = add_intercept ? 1 : 0;
  const int64_t n = Kf + ic;
  const bool want_vec = ic || b;
  void *dS, *dvec = nullptr, *dM, *dv = nullptr;
  PL_TRY(ws_get(WS_H_S, (size_t)Kf * Kf * 8, &dS));
  if (want_vec) PL_TRY(ws_get(WS_H_VEC, (size_t)(2 * Kf + 2) * 8, &dvec));
  PL_TRY(ws_get(WS_H_OUT1, (size_t)n * n * 8, &dM));
  if (AtWb) PL_TRY(ws_get(WS_H_OUT2, (size_t)n * 8, &dv));
  CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
  PL_TRY(syrk_dev_impl(dA, R, Kf, ld, dw, n_energy_rows, w_e, w_f, db, nullptr, (double*)dS, (double*)dvec, g.stream));
  PL_TRY(finish_dev_impl((double*)dS, (double*)dvec, Kf, ic, lambda, (double*)dM, (double*)dv, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
  CUDA_TRY(cudaMemcpyAsync(AtWA, dM, (size_t)n * n * 8, cudaMemcpyDeviceToHost, g.stream));
  if (AtWb) CUDA_TRY(cudaMemcpyAsync(AtWb, dv, (size_t)n * 8, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[3], g.stream));
  return finish_timing();
}

int pl_cov(const double* X, int64_t n, int64_t Kf, int64_t ldx, double* mean_inout, int mean_given, int center, double* C) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!X || !C || n <= 0 || Kf <= 0 || ldx < Kf) return set_err(PL_EINVAL, "bad arguments");
  if (center && !mean_inout) return set_err(PL_EINVAL, "centred covariance needs mean_inout");
  CUDA_TRY(cudaSetDevice(g.device));
  if (g_ndev > 1 && (double)n * (double)Kf * 8.0 >= MULTI_MIN_ROW_BYTES) return cov_multi(X, n, Kf, ldx, mean_inout, mean_given, center, C);
  CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
  double* dX;
  int64_t ld;
  PL_TRY(h2d_matrix(X, n, Kf, ldx, WS_H_IN1, &dX, &ld));
  double* dmean = nullptr;
  if (center) {
    if (mean_given) {
      PL_TRY(h2d_vector(mean_inout, Kf, WS_H_IN2, &dmean));
    } else {
      void* m;
      PL_TRY(ws_get(WS_H_IN2, (size_t)Kf * 8, &m));
      dmean = (double*)m;
    }
  }
  void *dS, *dC;
  PL_TRY(ws_get(WS_H_S, (size_t)Kf * Kf * 8, &dS));
  PL_TRY(ws_get(WS_H_OUT1, (size_t)Kf * Kf * 8, &dC));
  CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
  if (center && !mean_given) {
    // one pass over the rows: shift z = mean of a prefix, S_z and rho from the SYRK, then m = z + rho/n (pca.m = sum(d)/length(d),
    // pca_state.jl:35) and the re-centring about m
    void *z, *rho, *delta;
    PL_TRY(ws_get(WS_COV_Z, (size_t)Kf * 8, &z));
    PL_TRY(ws_get(WS_COV_RHO, (size_t)Kf * 8, &rho));
    PL_TRY(ws_get(WS_COV_DELTA, (size_t)Kf * 8, &delta));
    const int64_t prefix = n < 65536 ? n : 65536;
    PL_TRY(colsum(dX, prefix, Kf, ld, (double)prefix, (double*)z, g.stream));
    PL_TRY(cov_shifted_dev_impl(dX, n, Kf, ld, (const double*)z, (double*)dS, (double*)rho, g.stream));
    scale_div_kernel<<<(unsigned)((Kf + 255) / 256), 256, 0, g.stream>>>((const double*)rho, (double)n, Kf, (double*)delta);  // delta = m - z
    LAUNCH_CHECK();
    axpy1_kernel<<<(unsigned)((Kf + 255) / 256), 256, 0, g.stream>>>((const double*)z, (const double*)delta, Kf, dmean);      // m = z + delta
    LAUNCH_CHECK();
    PL_TRY(cov_recentre_dev_impl((double*)dS, Kf, (const double*)rho, (const double*)delta, (double)n, g.stream));
  } else {
    PL_TRY(syrk_dev_impl(dX, n, Kf, ld, nullptr, 0, 1.0, 1.0, nullptr, dmean, (double*)dS, nullptr, g.stream));
  }
  {
    const int64_t count = Kf * Kf;
    int blocks = (int)((count + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    scale_div_kernel<<<blocks, 256, 0, g.stream>>>((double*)dS, (double)n, count, (double*)dC);  // mean(di*di'): sum / n
    LAUNCH_CHECK();
  }
  CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
  CUDA_TRY(cudaMemcpyAsync(C, dC, (size_t)Kf * Kf * 8, cudaMemcpyDeviceToHost, g.stream));
  if (center && !mean_given) CUDA_TRY(cudaMemcpyAsync(mean_inout, dmean, (size_t)Kf * 8, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[3], g.stream));
  return finish_timing();
}


int pl_global_feature_dev(const double* dL, int64_t ldl, const int64_t* doffsets, int64_t N, int64_t d, int mean, double* dX, int64_t ldx,
                          int64_t total_rows, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  return global_feature_dev_impl(dL, ldl, doffsets, N, d, mean, dX, ldx, total_rows, (cudaStream_t)stream);
}

int pl_global_feature(const double* locals, int64_t ldl, const int64_t* offsets, int64_t N, int64_t d, int mean, double* X, int64_t ldx) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!locals || !offsets || !X || N < 0 || d <= 0 || ldl < d || ldx < d) return set_err(PL_EINVAL, "bad arguments");
  if (N == 0) return PL_OK;
  PL_TRY(check_offsets(offsets, N));
  const int64_t total = offsets[N];
  CUDA_TRY(cudaSetDevice(g.device));
  CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
  double* dL;
  int64_t ld;
  PL_TRY(h2d_matrix(locals, total, d, ldl, WS_FEAT_L, &dL, &ld));
  void *doff, *dX;
  PL_TRY(ws_get(WS_FEAT_OFF, (size_t)(N + 1) * 8, &doff));
  CUDA_TRY(cudaMemcpyAsync(doff, offsets, (size_t)(N + 1) * 8, cudaMemcpyHostToDevice, g.stream));
  const int64_t ldo = (d + 1) & ~(int64_t)1;
  PL_TRY(ws_get(WS_FEAT_X, (size_t)N * ldo * 8, &dX));
  CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
  PL_TRY(global_feature_dev_impl(dL, ld, (const int64_t*)doff, N, d, mean, (double*)dX, ldo, total, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
  CUDA_TRY(cudaMemcpy2DAsync(X, (size_t)ldx * 8, dX, (size_t)ldo * 8, (size_t)d * 8, (size_t)N, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[3], g.stream));
  return finish_timing();
}

int pl_normal_eq_blocks(const double* const* blocks, const int64_t* rows, int64_t nblocks, int64_t Kf, const double* w, int64_t n_energy_rows,
                        double w_e, double w_f, const double* b, int add_intercept, double lambda, int64_t chunk_rows, double* AtWA,
                        double* AtWb) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!blocks || !rows || !AtWA || nblocks < 0 || Kf <= 0) return set_err(PL_EINVAL, "bad arguments");
  if (AtWb && !b) return set_err(PL_EINVAL, "AtWb requested without b");
  int64_t R = 0;
  for (int64_t k = 0; k < nblocks; ++k) {
    if (rows[k] < 0 || (rows[k] > 0 && !blocks[k])) return set_err(PL_EINVAL, "block %lld: bad row count or NULL pointer", (long long)k);
    R += rows[k];
  }
  CUDA_TRY(cudaSetDevice(g.device));
  if (chunk_rows <= 0) {
    chunk_rows = ((int64_t)512 << 20) / (Kf * 8);  // ~512 MB per buffer: long enough copies for full PCIe rate, SYRK still > 1 wave
    if (chunk_rows < 4096) chunk_rows = 4096;
  }
  chunk_rows = (chunk_rows + 15) / 16 * 16;
  // dense blocks: leading dimension Kf each
  int64_t* lds = (int64_t*)malloc((size_t)(nblocks > 0 ? nblocks : 1) * sizeof(int64_t));
  if (!lds) return set_err(PL_ENOMEM, "host allocation failed");
  for (int64_t k = 0; k < nblocks; ++k) lds[k] = Kf;
  int64_t ne = n_energy_rows < 0 ? 0 : (n_energy_rows > R ? R : n_energy_rows);
  const int rc = (g_ndev > 1 && (double)R * (double)Kf * 8.0 >= MULTI_MIN_ROW_BYTES)
                     ? normal_eq_multi(blocks, rows, lds, nblocks, R, Kf, w, ne, w_e, w_f, b, add_intercept ? 1 : 0, lambda, AtWA, AtWb, chunk_rows)
                     : normal_eq_stream(blocks, rows, lds, nblocks, R, Kf, w, ne, w_e, w_f, b, add_intercept ? 1 : 0, lambda, AtWA, AtWb, chunk_rows);
  free(lds);
  return rc;
}


int pl_eigh_dev(double* dA, int64_t n, int64_t lda, int fill, int want_vectors, double* dW, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  return eigh_dev_impl(dA, n, lda, fill, want_vectors, dW, (cudaStream_t)stream);
}

int pl_eigh(const double* A, int64_t n, int64_t lda, int fill, double* W, double* V, int64_t ldv) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!A || !W || n <= 0 || lda < n || (V && ldv < n)) return set_err(PL_EINVAL, "bad arguments");
  CUDA_TRY(cudaSetDevice(g.device));
  CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
  void *dA, *dW;  // not the WS_EIG_* slots: a resident L-ensemble survives plain eigendecompositions
  PL_TRY(ws_get(WS_H_OUT1, (size_t)n * n * 8, &dA));
  PL_TRY(ws_get(WS_H_OUT2, (size_t)n * 8, &dW));
  CUDA_TRY(cudaMemcpy2DAsync(dA, (size_t)n * 8, A, (size_t)lda * 8, (size_t)n * 8, (size_t)n, cudaMemcpyHostToDevice, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
  PL_TRY(eigh_dev_impl((double*)dA, n, n, fill, V != nullptr, (double*)dW, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
  CUDA_TRY(cudaMemcpyAsync(W, dW, (size_t)n * 8, cudaMemcpyDeviceToHost, g.stream));
  if (V) CUDA_TRY(cudaMemcpy2DAsync(V, (size_t)ldv * 8, dA, (size_t)n * 8, (size_t)n * 8, (size_t)n, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[3], g.stream));
  return finish_timing();
}

int pl_ell_ensemble(int kernel_id, const double* X, int64_t n, int64_t d, int64_t ldx, int alpha, int cinv_kind, const double* cinv, double ell,
                    double beta, double* lambda, double* Kout, int64_t ldk) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!X || !lambda || n <= 0 || d <= 0 || ldx < d || (Kout && ldk < n)) return set_err(PL_EINVAL, "bad arguments");
  CUDA_TRY(cudaSetDevice(g.device));
  ell_state.valid = false;
  CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
  double *dX, *dc = nullptr;
  int64_t ld;
  PL_TRY(h2d_matrix(X, n, d, ldx, WS_H_IN1, &dX, &ld));
  if ((kernel_id == PL_KERNEL_RBF || kernel_id == PL_KERNEL_IMQ) && cinv_kind != PL_CINV_IDENTITY) {
    if (!cinv) return set_err(PL_EINVAL, "cinv is NULL");
    PL_TRY(h2d_vector(cinv, cinv_kind == PL_CINV_DIAGONAL ? d : d * d, WS_H_IN3, &dc));
  }
  const int64_t ldo = (n + 1) & ~(int64_t)1;
  void *dA, *dW;
  PL_TRY(ws_get(WS_EIG_A, (size_t)n * ldo * 8, &dA));
  PL_TRY(ws_get(WS_EIG_W, (size_t)n * 8, &dW));
  CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
  // K = KernelMatrix(F, k) (dpp.jl:25,40) in the Symmetric(:U) storage, straight into the eigensolver's buffer
  PL_TRY(gram_dev_impl(kernel_id, dX, n, ld, nullptr, 0, 0, d, alpha, cinv_kind, dc, ell, beta, (double*)dA, ldo, PL_FILL_JULIA_U, 0, 1, 0, g.stream));
  // a copy of the (triangle) storage is kept for pl_ell_greedy, which needs rows of K, up to n = 32768 (8.6 GB)
  ell_state.has_kernel = false;
  if (n <= 32768) {
    void* dKk;
    PL_TRY(ws_get(WS_ELL_K, (size_t)n * ldo * 8, &dKk));
    CUDA_TRY(cudaMemcpyAsync(dKk, dA, (size_t)n * ldo * 8, cudaMemcpyDeviceToDevice, g.stream));
    ell_state.has_kernel = true;
  }
  if (Kout) {  // the caller also wants K itself (kDPP keeps it as ell.L): triangle rows, before the solver overwrites the buffer
    const int64_t edge = gram_small_tiles(n, n, true, 1, kernel_id) ? 64 : GRAM_BM;
    for (int64_t r0 = 0; r0 < n; r0 += edge) {
      const int64_t r1 = r0 + edge < n ? r0 + edge : n;
      CUDA_TRY(cudaMemcpy2DAsync(Kout + r0 * ldk, (size_t)ldk * 8, (double*)dA + r0 * ldo, (size_t)ldo * 8, (size_t)r1 * 8, (size_t)(r1 - r0), cudaMemcpyDeviceToHost, g.stream));
    }
  }
  PL_TRY(eigh_dev_impl((double*)dA, n, ldo, PL_FILL_JULIA_U, 1, (double*)dW, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
  CUDA_TRY(cudaMemcpyAsync(lambda, dW, (size_t)n * 8, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[3], g.stream));
  PL_TRY(finish_timing());
  ell_state.n = n;
  ell_state.ld = ldo;
  ell_state.valid = true;
  return PL_OK;
}

int pl_ell_inclusion_prob(double alpha_scale, int64_t n, double* p) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!ell_state.valid) return set_err(PL_EINVAL, "no resident L-ensemble: call pl_ell_ensemble first");
  if (n != ell_state.n || !p) return set_err(PL_EINVAL, "bad arguments (resident ensemble has %lld items)", (long long)ell_state.n);
  CUDA_TRY(cudaSetDevice(g.device));
  void *dg, *dp;
  PL_TRY(ws_get(WS_EIG_GAMMA, (size_t)n * 8, &dg));
  PL_TRY(ws_get(WS_EIG_P, (size_t)n * 8, &dp));
  CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
  ell_gamma_kernel<<<(unsigned)((n + 255) / 256), 256, 0, g.stream>>>((const double*)g.ws[WS_EIG_W], n, alpha_scale, (double*)dg);
  LAUNCH_CHECK();
  PL_TRY(inclusion_prob_dev_impl((const double*)g.ws[WS_EIG_A], n, n, ell_state.ld, (const double*)dg, (double*)dp, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
  CUDA_TRY(cudaMemcpyAsync(p, dp, (size_t)n * 8, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[3], g.stream));
  return finish_timing();
}

int pl_ell_eigvecs(const int64_t* idx, int64_t m, int64_t n, double* V, int64_t ldv) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!ell_state.valid) return set_err(PL_EINVAL, "no resident L-ensemble: call pl_ell_ensemble first");
  if (n != ell_state.n || m < 0 || (m > 0 && (!idx || !V)) || ldv < n) return set_err(PL_EINVAL, "bad arguments");
  if (m == 0) return PL_OK;
  for (int64_t r = 0; r < m; ++r)
    if (idx[r] < 0 || idx[r] >= n) return set_err(PL_EINVAL, "eigenvector index %lld out of range", (long long)idx[r]);
  CUDA_TRY(cudaSetDevice(g.device));
  void *didx, *dsel;
  PL_TRY(ws_get(WS_EIG_IDX, (size_t)m * 8, &didx));
  PL_TRY(ws_get(WS_EIG_SEL, (size_t)m * n * 8, &dsel));
  CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
  CUDA_TRY(cudaMemcpyAsync(didx, idx, (size_t)m * 8, cudaMemcpyHostToDevice, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
  int64_t rem = m, done = 0;
  while (rem > 0) {  // grid.y is limited to 65535
    const int64_t take = rem < 32768 ? rem : 32768;
    dim3 grid((unsigned)((n + 1023) / 1024 < 8 ? (n + 1023) / 1024 : 8), (unsigned)take);
    gather_rows_kernel<<<grid, 256, 0, g.stream>>>((const double*)g.ws[WS_EIG_A], n, ell_state.ld, (const int64_t*)didx + done, take, (double*)dsel + done * n);
    LAUNCH_CHECK();
    rem -= take;
    done += take;
  }
  CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
  CUDA_TRY(cudaMemcpy2DAsync(V, (size_t)ldv * 8, dsel, (size_t)n * 8, (size_t)n * 8, (size_t)m, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[3], g.stream));
  return finish_timing();
}


// ---- out-of-core accumulation: ooc_learn! (linear-learn.jl:301-391) keeps AtWA / AtWb and adds one configuration at a time --------------
int pl_neq_begin(int64_t Kf) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (Kf <= 0) return set_err(PL_EINVAL, "bad descriptor length %lld", (long long)Kf);
  CUDA_TRY(cudaSetDevice(g.device));
  void *dS, *dV;
  PL_TRY(ws_get(WS_ACC_S, (size_t)Kf * Kf * 8, &dS));
  PL_TRY(ws_get(WS_ACC_V, (size_t)(2 * Kf + 2) * 8, &dV));
  CUDA_TRY(cudaMemsetAsync(dS, 0, (size_t)Kf * Kf * 8, g.stream));
  CUDA_TRY(cudaMemsetAsync(dV, 0, (size_t)(2 * Kf + 2) * 8, g.stream));
  neq_state.Kf = Kf;
  neq_state.rows = 0;
  neq_state.batches = 0;
  neq_state.open = true;
  return PL_OK;
}

int pl_neq_add(const double* A, int64_t R, int64_t lda, const double* w, int64_t n_energy_rows, double w_e, double w_f, const double* b) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!neq_state.open) return set_err(PL_EINVAL, "pl_neq_add without pl_neq_begin");
  const int64_t Kf = neq_state.Kf;
  if (!A || R < 0 || lda < Kf) return set_err(PL_EINVAL, "bad arguments");
  if (R == 0) return PL_OK;
  CUDA_TRY(cudaSetDevice(g.device));
  CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
  double *dA, *dw = nullptr, *db = nullptr;
  int64_t ld;
  PL_TRY(h2d_matrix(A, R, Kf, lda, WS_H_IN1, &dA, &ld));
  if (w) PL_TRY(h2d_vector(w, R, WS_H_IN2, &dw));
  if (b) PL_TRY(h2d_vector(b, R, WS_H_IN3, &db));
  void *dS, *dvec;
  PL_TRY(ws_get(WS_H_S, (size_t)Kf * Kf * 8, &dS));
  PL_TRY(ws_get(WS_H_VEC, (size_t)(2 * Kf + 2) * 8, &dvec));
  CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
  const int64_t ne = n_energy_rows < 0 ? 0 : (n_energy_rows > R ? R : n_energy_rows);
  PL_TRY(syrk_dev_impl(dA, R, Kf, ld, dw, ne, w_e, w_f, db, nullptr, (double*)dS, (double*)dvec, g.stream));
  accumulate_kernel<<<g.num_sms * 4, 256, 0, g.stream>>>((const double*)dS, Kf * Kf, (double*)g.ws[WS_ACC_S]);
  LAUNCH_CHECK();
  accumulate_kernel<<<8, 256, 0, g.stream>>>((const double*)dvec, 2 * Kf + 2, (double*)g.ws[WS_ACC_V]);
  LAUNCH_CHECK();
  CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[3], g.stream));
  neq_state.rows += R;
  neq_state.batches += 1;
  return finish_timing();  // synchronises: the host buffers of this batch may be reused by the caller
}

int pl_neq_finish(int add_intercept, double lambda, double* AtWA, double* AtWb) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!neq_state.open) return set_err(PL_EINVAL, "pl_neq_finish without pl_neq_begin");
  if (!AtWA) return set_err(PL_EINVAL, "null pointer");
  const int64_t Kf = neq_state.Kf;
  const int ic = add_intercept ? 1 : 0;
  const int64_t n = Kf + ic;
  CUDA_TRY(cudaSetDevice(g.device));
  void *dM, *dv = nullptr;
  PL_TRY(ws_get(WS_H_OUT1, (size_t)n * n * 8, &dM));
  if (AtWb) PL_TRY(ws_get(WS_H_OUT2, (size_t)n * 8, &dv));
  CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
  PL_TRY(finish_dev_impl((const double*)g.ws[WS_ACC_S], (const double*)g.ws[WS_ACC_V], Kf, ic, lambda, (double*)dM, (double*)dv, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
  CUDA_TRY(cudaMemcpyAsync(AtWA, dM, (size_t)n * n * 8, cudaMemcpyDeviceToHost, g.stream));
  if (AtWb) CUDA_TRY(cudaMemcpyAsync(AtWb, dv, (size_t)n * 8, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[3], g.stream));
  return finish_timing();  // the accumulators stay open: more batches may follow (ooc_learn! can be resumed with AtWA / AtWb)
}


int pl_correlation_feature(const double* locals, int64_t ldl, const int64_t* offsets, int64_t N, int64_t d, double* C) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!locals || !offsets || !C || N < 0 || d <= 0 || ldl < d) return set_err(PL_EINVAL, "bad arguments");
  if (N == 0) return PL_OK;
  PL_TRY(check_offsets(offsets, N));
  const int64_t total = offsets[N];
  CUDA_TRY(cudaSetDevice(g.device));
  CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
  double* dL;
  int64_t ld;
  PL_TRY(h2d_matrix(locals, total, d, ldl, WS_FEAT_L, &dL, &ld));
  void *doff, *dC;
  PL_TRY(ws_get(WS_FEAT_OFF, (size_t)(N + 1) * 8, &doff));
  CUDA_TRY(cudaMemcpyAsync(doff, offsets, (size_t)(N + 1) * 8, cudaMemcpyHostToDevice, g.stream));
  PL_TRY(ws_get(WS_FEAT_X, (size_t)N * d * d * 8, &dC));
  CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
  correlation_feature_kernel<<<(unsigned)N, 256, 0, g.stream>>>(dL, ld, (const int64_t*)doff, d, (double*)dC);
  LAUNCH_CHECK();
  CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
  CUDA_TRY(cudaMemcpyAsync(C, dC, (size_t)N * d * d * 8, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[3], g.stream));
  return finish_timing();
}


int pl_ksd_rbf(const double* X, const double* S, int64_t n, int64_t d, int64_t ldx, int64_t lds, int cinv_kind, const double* cinv, double ell,
               double beta, double* ksd) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!X || !S || !ksd || n < 2 || d <= 0 || ldx < d || lds < d) return set_err(PL_EINVAL, "bad arguments (KSD needs n >= 2 samples)");
  if (n > 65535) return set_err(PL_EINVAL, "KSD supports up to 65535 samples (four n x n Gram matrices are kept on the device)");
  if (!(ell != 0.0)) return set_err(PL_EINVAL, "length scale must be non-zero");
  if (cinv_kind != PL_CINV_IDENTITY && cinv_kind != PL_CINV_DIAGONAL && cinv_kind != PL_CINV_FULL) return set_err(PL_EINVAL, "unknown cinv_kind %d", cinv_kind);
  if (cinv_kind != PL_CINV_IDENTITY && !cinv) return set_err(PL_EINVAL, "cinv is NULL");
  CUDA_TRY(cudaSetDevice(g.device));
  cudaStream_t st = g.stream;
  CUDA_TRY(cudaEventRecord(g.ev[0], st));
  double *dX, *dS, *dc = nullptr;
  int64_t l1, l2;
  PL_TRY(h2d_matrix(X, n, d, ldx, WS_H_IN1, &dX, &l1));
  PL_TRY(h2d_matrix(S, n, d, lds, WS_KSD_S, &dS, &l2));
  double trC = (double)d;
  if (cinv_kind == PL_CINV_DIAGONAL) {
    PL_TRY(h2d_vector(cinv, d, WS_H_IN3, &dc));
    trC = 0.0;
    for (int64_t k = 0; k < d; ++k) trC += cinv[k];
  } else if (cinv_kind == PL_CINV_FULL) {
    PL_TRY(h2d_vector(cinv, d * d, WS_H_IN3, &dc));
    trC = 0.0;
    for (int64_t k = 0; k < d; ++k) trC += cinv[k * d + k];
  }
  CUDA_TRY(cudaEventRecord(g.ev[1], st));
  // Xc = X - mean (the pair differences are translation invariant; the Gram forms below must not cancel against |mean|^2)
  const int64_t ldc = (d + 1) & ~(int64_t)1;
  void *mean, *xc, *y = nullptr;
  PL_TRY(ws_get(WS_MEAN, (size_t)ldc * 8, &mean));
  PL_TRY(colsum(dX, n, d, l1, (double)n, (double*)mean, st));
  PL_TRY(ws_get(WS_XC1, (size_t)n * ldc * 8, &xc));
  PL_TRY(launch_prep<PREP_CENTER_ONLY>(dX, n, d, l1, (double*)mean, nullptr, (double*)xc, nullptr, ldc, nullptr, st));
  const double* Y = (const double*)xc;
  if (cinv_kind != PL_CINV_IDENTITY) {
    PL_TRY(ws_get(WS_Y1, (size_t)n * ldc * 8, &y));
    if (cinv_kind == PL_CINV_DIAGONAL) {
      int blocks = (int)((n * d + 255) / 256 < g.num_sms * 8 ? (n * d + 255) / 256 : g.num_sms * 8);
      scale_cols_kernel<<<blocks, 256, 0, st>>>((const double*)xc, n, d, ldc, dc, (double*)y);
      LAUNCH_CHECK();
    } else {
      const double* cf;
      int64_t ldcf;
      PL_TRY(tma_operand(dc, d, d, d, WS_PADC, st, &cf, &ldcf));
      GramLaunch Ly{};
      Ly.epi = EPI_LINEAR;
      Ly.A = (const double*)xc; Ly.nA = n; Ly.ldA = ldc;
      Ly.B = cf; Ly.nB = d; Ly.ldB = ldcf;
      Ly.d = d;
      Ly.out = (double*)y; Ly.ldo = ldc;
      Ly.nparts = 1;
      PL_TRY(launch_gram(Ly, st));
    }
    Y = (const double*)y;
  }
  const int64_t ldg = (n + 1) & ~(int64_t)1;
  void *g1, *g2, *g3, *h;
  PL_TRY(ws_get(WS_KSD_G2, (size_t)n * ldg * 8, &g2));
  PL_TRY(ws_get(WS_KSD_G3, (size_t)n * ldg * 8, &g3));
  PL_TRY(ws_get(WS_KSD_H, (size_t)n * ldg * 8, &h));
  auto linear = [&](const double* A, int64_t lda, const double* B, int64_t ldb, bool sym, double* out) -> int {
    GramLaunch L{};
    L.epi = EPI_LINEAR;
    L.A = A; L.nA = n; L.ldA = lda;
    L.B = B; L.nB = n; L.ldB = ldb;
    L.d = d;
    L.out = out; L.ldo = ldg;
    L.symmetric = sym ? 1 : 0;
    L.same_operand = sym ? 1 : 0;
    L.fill = PL_FILL_FULL;
    L.nparts = 1;
    return launch_gram(L, st);
  };
  PL_TRY(linear(Y, ldc, Y, ldc, true, (double*)g2));                       // G2 = Y Y'
  PL_TRY(linear(dS, l2, dS, l2, true, (double*)g3));                      // G3 = S S'
  PL_TRY(linear(dS, l2, Y, ldc, false, (double*)h));                      // H  = S Y'
  if (cinv_kind == PL_CINV_IDENTITY) {
    g1 = g2;                                                               // Y == Xc
  } else {
    PL_TRY(ws_get(WS_KSD_G1, (size_t)n * ldg * 8, &g1));
    PL_TRY(linear((const double*)xc, ldc, Y, ldc, false, (double*)g1));   // G1 = Xc Y' (symmetric because Cinv is)
  }
  int bx = (int)((n + 255) / 256);
  if (bx > 8) bx = 8;
  void *part, *dout;
  PL_TRY(ws_get(WS_KSD_PART, (size_t)(n * bx + 1) * 8, &part));
  dout = (double*)part + n * bx;
  ksd_reduce_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, st>>>((const double*)g1, (const double*)g2, (const double*)g3, (const double*)h, n, ldg, beta,
                                                                     1.0 / (ell * ell), trC, (double)d, (double*)part);
  LAUNCH_CHECK();
  ksd_final_kernel<<<1, 1024, 0, st>>>((const double*)part, n * bx, (double)n * ((double)n - 1.0), (double*)dout);
  LAUNCH_CHECK();
  CUDA_TRY(cudaEventRecord(g.ev[2], st));
  CUDA_TRY(cudaMemcpyAsync(ksd, dout, 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(g.ev[3], st));
  return finish_timing();
}


int pl_cov_shifted_dev(const double* dA, int64_t R, int64_t Kf, int64_t lda, const double* dz, double* dS, double* drho, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  return cov_shifted_dev_impl(dA, R, Kf, lda, dz, dS, drho, (cudaStream_t)stream);
}

int pl_cov_recentre_dev(double* dS, int64_t Kf, const double* drho, const double* ddelta, double n_rows, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  return cov_recentre_dev_impl(dS, Kf, drho, ddelta, n_rows, (cudaStream_t)stream);
}


// Pinned (page-locked), zero-filled host memory for result arrays: D2H into pinned memory runs at the PCIe rate (config 2: 37 ms end to
// end); pageable memory goes through the staging ring (45 ms; ~100 ms when the pages have never been touched).
void* pl_host_alloc(int64_t bytes) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (need_ready() != PL_OK || bytes <= 0) return nullptr;
  void* p = nullptr;
  if (cudaSetDevice(g.device) != cudaSuccess || cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    set_err(PL_ENOMEM, "pinned host allocation of %lld bytes failed", (long long)bytes);
    return nullptr;
  }
  memset(p, 0, (size_t)bytes);  // zeros(n, n) (kernels.jl:216)
  return p;
}

int pl_host_free(void* p) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!p) return PL_OK;
  CUDA_TRY(cudaFreeHost(p));
  return PL_OK;
}


// First phase of a k-DPP draw (Kulesza & Taskar 2012 Alg. 8; Determinantal.sample, dpp.jl:51 [external]): which eigenvectors.
int pl_esp_select(const double* lam, int64_t n, int64_t k, const double* uniforms, int64_t* idx, int64_t* count) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!lam || !uniforms || !idx || !count || n <= 0 || k < 0 || k > n) return set_err(PL_EINVAL, "bad arguments");
  *count = 0;
  if (k == 0) return PL_OK;
  int64_t positive = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (!(lam[i] >= 0.0) || !std::isfinite(lam[i])) return set_err(PL_EINVAL, "eigenvalue %lld is negative or not finite", (long long)i);
    if (!(uniforms[i] >= 0.0 && uniforms[i] < 1.0)) return set_err(PL_EINVAL, "uniforms must lie in [0, 1)");
    positive += lam[i] > 0.0;
  }
  if (k > positive) return set_err(PL_EINVAL, "k = %lld exceeds the number of positive eigenvalues (%lld)", (long long)k, (long long)positive);
  const int64_t ldt = (k + 2) & ~(int64_t)1;
  const double table_gib = (double)(n + 1) * (double)ldt * 8.0 / 1073741824.0;
  if (table_gib > 24.0) return set_err(PL_ENOMEM, "the (n+1) x (k+1) table of elementary symmetric polynomials needs %.1f GiB (limit 24)", table_gib);
  CUDA_TRY(cudaSetDevice(g.device));
  cudaStream_t st = g.stream;
  void *dT, *daux;
  PL_TRY(ws_get(WS_ESP_T, (size_t)(n + 1) * ldt * 8, &dT));
  PL_TRY(ws_get(WS_ESP_AUX, (size_t)(3 * n + k + 2) * 8, &daux));
  double* dlam = (double*)daux;
  double* dll = dlam + n;
  double* du = dll + n;
  int64_t* didx = (int64_t*)(du + n);
  int64_t* dcnt = didx + k;
  CUDA_TRY(cudaEventRecord(g.ev[0], st));
  CUDA_TRY(cudaMemcpyAsync(dlam, lam, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(du, uniforms, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaEventRecord(g.ev[1], st));
  esp_loglam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dlam, n, dll);
  LAUNCH_CHECK();
  esp_forward_kernel<<<ESP_CLUSTER, 1024, 0, st>>>(dll, n, k, (double*)dT, ldt);
  LAUNCH_CHECK();
  esp_select_kernel<<<1, 32, 0, st>>>((const double*)dT, ldt, dll, n, k, du, didx, dcnt);
  LAUNCH_CHECK();
  CUDA_TRY(cudaEventRecord(g.ev[2], st));
  int64_t cnt = 0;
  CUDA_TRY(cudaMemcpyAsync(&cnt, dcnt, 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(g.ev[3], st));
  PL_TRY(finish_timing());
  if (cnt < 0 || cnt > k) return set_err(PL_ECUDA, "selection pass returned %lld eigenvectors for k = %lld", (long long)cnt, (long long)k);
  if (cnt > 0) CUDA_TRY(cudaMemcpy(idx, didx, (size_t)cnt * 8, cudaMemcpyDeviceToHost));
  *count = cnt;
  return PL_OK;
}

// pl_ell_sample in its kernel-row (Cholesky) form, see prelude.cuh: P = V V' once on the DMMA engine, then per step
// pick -> gather C[:t,item] -> C[:t,:]'cj (two-stage, deterministic) -> row update.
static int ell_sample_chol(const int64_t* eig_idx, int64_t m, int64_t n, const double* uniforms, int64_t* items, cudaStream_t st) {
  const int64_t ldv = (m + 1) & ~(int64_t)1, ldp = (n + 1) & ~(int64_t)1;
  void *didx, *dV, *dP, *dC, *dp, *dvec, *dtmp, *ditems, *du;
  PL_TRY(ws_get(WS_EIG_IDX, (size_t)m * 8, &didx));
  PL_TRY(ws_get(WS_EIG_SEL, (size_t)n * ldv * 8, &dV));
  PL_TRY(ws_get(WS_DPP_K, (size_t)n * ldp * 8, &dP));
  PL_TRY(ws_get(WS_DPP_C, (size_t)m * n * 8, &dC));
  PL_TRY(ws_get(WS_DPP_P, (size_t)n * 8, &dp));
  PL_TRY(ws_get(WS_DPP_VEC, (size_t)(m + 2) * 8, &dvec));
  int nslices = (int)((m + 255) / 256);
  if (nslices > 16) nslices = 16;
  PL_TRY(ws_get(WS_DPP_TMP, (size_t)nslices * n * 8, &dtmp));
  PL_TRY(ws_get(WS_DPP_ITEMS, (size_t)m * 8, &ditems));
  PL_TRY(ws_get(WS_DPP_U, (size_t)m * 8, &du));
  double* cj = (double*)dvec;
  double* pivot = cj + m;
  CUDA_TRY(cudaEventRecord(g.ev[0], st));
  CUDA_TRY(cudaMemcpyAsync(didx, eig_idx, (size_t)m * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(du, uniforms, (size_t)m * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(ditems, 0xff, (size_t)m * 8, st));  // -1: "no item picked" stays visible
  CUDA_TRY(cudaEventRecord(g.ev[1], st));
  {
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((ldv + 31) / 32));
    gather_cols_kernel<<<grid, dim3(32, 8), 0, st>>>((const double*)g.ws[WS_EIG_A], n, ell_state.ld, (const int64_t*)didx, m, (double*)dV, ldv);
    LAUNCH_CHECK();
    GramLaunch L{};
    L.epi = EPI_LINEAR;
    L.A = (const double*)dV; L.nA = n; L.ldA = ldv;
    L.B = (const double*)dV; L.nB = n; L.ldB = ldv;
    L.d = m;
    L.out = (double*)dP; L.ldo = ldp;
    L.symmetric = 1;
    L.same_operand = 1;
    L.fill = PL_FILL_FULL;
    L.nparts = 1;
    PL_TRY(launch_gram(L, st));
  }
  const unsigned nb = (unsigned)((n + 127) / 128);
  greedy_init_kernel<<<nb, 128, 0, st>>>((const double*)dP, n, ldp, (double*)dp);
  LAUNCH_CHECK();
  for (int64_t t = 0; t < m; ++t) {
    dpp_pick_kernel<<<1, 1024, 0, st>>>((const double*)dp, n, (const double*)du, t, (int64_t*)ditems);
    LAUNCH_CHECK();
    dpp_chol_gather_kernel<<<(unsigned)(t > 0 ? (t + 127) / 128 : 1), 128, 0, st>>>((const double*)dC, t, n, (const int64_t*)ditems, t, (const double*)dp, cj, pivot);
    LAUNCH_CHECK();
    int sl = 0;
    if (t > 0) {
      sl = (int)((t + 255) / 256);
      if (sl > nslices) sl = nslices;
      const int64_t sslice = (t + sl - 1) / sl;
      greedy_dot_partial_kernel<<<dim3(nb, (unsigned)sl), 128, 0, st>>>((const double*)dC, t, n, cj, sslice, (double*)dtmp);
      LAUNCH_CHECK();
    }
    dpp_chol_finish_kernel<<<nb, 128, 0, st>>>((const double*)dP, n, ldp, (const double*)dtmp, sl, (const int64_t*)ditems, t, pivot, (double*)dC + t * n, (double*)dp);
    LAUNCH_CHECK();
  }
  CUDA_TRY(cudaEventRecord(g.ev[2], st));
  CUDA_TRY(cudaMemcpyAsync(items, ditems, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(g.ev[3], st));
  PL_TRY(finish_timing());
  for (int64_t r = 0; r < m; ++r)
    if (items[r] < 0) return set_err(PL_ECUDA, "k-DPP draw: step %lld found no admissible item (the selected eigenvectors are rank deficient)", (long long)r);
  return PL_OK;
}

int pl_ell_sample(const int64_t* eig_idx, int64_t m, int64_t n, const double* uniforms, int64_t* items) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!ell_state.valid) return set_err(PL_EINVAL, "no resident L-ensemble: call pl_ell_ensemble first");
  if (n != ell_state.n || m < 0 || m > n || (m > 0 && (!eig_idx || !uniforms || !items))) return set_err(PL_EINVAL, "bad arguments");
  if (m == 0) return PL_OK;
  for (int64_t r = 0; r < m; ++r) {
    if (eig_idx[r] < 0 || eig_idx[r] >= n) return set_err(PL_EINVAL, "eigenvector index %lld out of range", (long long)eig_idx[r]);
    if (!(uniforms[r] >= 0.0 && uniforms[r] < 1.0)) return set_err(PL_EINVAL, "uniforms must lie in [0, 1)");
  }
  CUDA_TRY(cudaSetDevice(g.device));
  cudaStream_t st = g.stream;
  // n x n projection kernel affordable (<= 8.6 GB): the kernel-row form, one pass over the rows of C per step
  if (n <= 32768 && !getenv("PLB200_DPP_GS")) return ell_sample_chol(eig_idx, m, n, uniforms, items, st);
  void *didx, *dVt, *dp, *dF, *dvec, *dtmp, *ditems, *du;
  PL_TRY(ws_get(WS_EIG_IDX, (size_t)m * 8, &didx));
  PL_TRY(ws_get(WS_EIG_SEL, (size_t)m * n * 8, &dVt));
  PL_TRY(ws_get(WS_DPP_P, (size_t)n * 8, &dp));
  PL_TRY(ws_get(WS_DPP_F, (size_t)m * m * 8, &dF));
  PL_TRY(ws_get(WS_DPP_VEC, (size_t)2 * m * 8, &dvec));
  int nslices = (int)((m + 511) / 512);  // enough blocks in flight for the pass over Vt: (n / 128) x nslices
  if (nslices > 16) nslices = 16;
  const int64_t kslice = (m + nslices - 1) / nslices;
  PL_TRY(ws_get(WS_DPP_TMP, (size_t)nslices * n * 8, &dtmp));
  PL_TRY(ws_get(WS_DPP_ITEMS, (size_t)m * 8, &ditems));
  PL_TRY(ws_get(WS_DPP_U, (size_t)m * 8, &du));
  double* v = (double*)dvec;
  double* c = v + m;
  CUDA_TRY(cudaEventRecord(g.ev[0], st));
  CUDA_TRY(cudaMemcpyAsync(didx, eig_idx, (size_t)m * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(du, uniforms, (size_t)m * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(ditems, 0xff, (size_t)m * 8, st));  // -1: "no item picked" stays visible
  CUDA_TRY(cudaEventRecord(g.ev[1], st));
  {  // Vt = the selected rows of the resident eigenvector matrix (m x n, dense)
    int64_t rem = m, done = 0;
    while (rem > 0) {
      const int64_t take = rem < 32768 ? rem : 32768;
      dim3 grid((unsigned)((n + 1023) / 1024 < 8 ? (n + 1023) / 1024 : 8), (unsigned)take);
      gather_rows_kernel<<<grid, 256, 0, st>>>((const double*)g.ws[WS_EIG_A], n, ell_state.ld, (const int64_t*)didx + done, take, (double*)dVt + done * n);
      LAUNCH_CHECK();
      rem -= take;
      done += take;
    }
  }
  const unsigned nb = (unsigned)((n + 127) / 128), mb = (unsigned)((m + 127) / 128);
  dpp_init_p_kernel<<<nb, 128, 0, st>>>((const double*)dVt, m, n, (double*)dp);
  LAUNCH_CHECK();
  for (int64_t i = 0; i < m; ++i) {
    dpp_pick_kernel<<<1, 1024, 0, st>>>((const double*)dp, n, (const double*)du, i, (int64_t*)ditems);
    LAUNCH_CHECK();
    dpp_gather_v_kernel<<<mb, 128, 0, st>>>((const double*)dVt, m, n, (const int64_t*)ditems, i, v);
    LAUNCH_CHECK();
    if (i > 0) {
      for (int pass = 0; pass < 2; ++pass) {  // classical Gram-Schmidt, twice (keeps F orthonormal to rounding for thousands of steps)
        dpp_rowdots_kernel<<<(unsigned)((i * 32 + 255) / 256), 256, 0, st>>>((const double*)dF, i, m, v, c);
        LAUNCH_CHECK();
        int jsl = (int)((i + 255) / 256);  // row slices of >= 256 rows, at most as many as the Vt pass uses (tmp is nslices x n >= jsl x m)
        if (jsl > nslices) jsl = nslices;
        if (jsl < 1) jsl = 1;
        const int64_t jslice = (i + jsl - 1) / jsl;
        dpp_subtract_partial_kernel<<<dim3(mb, (unsigned)jsl), 128, 0, st>>>((const double*)dF, i, m, c, jslice, (double*)dtmp);
        LAUNCH_CHECK();
        dpp_subtract_finish_kernel<<<mb, 128, 0, st>>>((const double*)dtmp, jsl, m, v);
        LAUNCH_CHECK();
      }
    }
    dpp_normalise_kernel<<<1, 1024, 0, st>>>(v, m, (double*)dF + i * m);
    LAUNCH_CHECK();
    dpp_update_partial_kernel<<<dim3(nb, (unsigned)nslices), 128, 0, st>>>((const double*)dVt, m, n, (const double*)dF + i * m, kslice, (double*)dtmp);
    LAUNCH_CHECK();
    dpp_update_finish_kernel<<<nb, 128, 0, st>>>((const double*)dtmp, nslices, n, (const int64_t*)ditems, i, (double*)dp);
    LAUNCH_CHECK();
  }
  CUDA_TRY(cudaEventRecord(g.ev[2], st));
  CUDA_TRY(cudaMemcpyAsync(items, ditems, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(g.ev[3], st));
  PL_TRY(finish_timing());
  for (int64_t r = 0; r < m; ++r)
    if (items[r] < 0) return set_err(PL_ECUDA, "k-DPP draw: step %lld found no admissible item (the selected eigenvectors are rank deficient)", (long long)r);
  return PL_OK;
}


int pl_ell_greedy(int64_t k, int64_t n, int64_t* items) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!ell_state.valid) return set_err(PL_EINVAL, "no resident L-ensemble: call pl_ell_ensemble first");
  if (!ell_state.has_kernel) return set_err(PL_EINVAL, "the kernel matrix of this ensemble was not kept on the device (n > 32768)");
  if (n != ell_state.n || k < 0 || k > n || (k > 0 && !items)) return set_err(PL_EINVAL, "bad arguments");
  if (k == 0) return PL_OK;
  CUDA_TRY(cudaSetDevice(g.device));
  cudaStream_t st = g.stream;
  const double* K = (const double*)g.ws[WS_ELL_K];
  const int64_t ld = ell_state.ld;
  void *dd, *dC, *dcj, *dtmp, *ditems;
  PL_TRY(ws_get(WS_DPP_P, (size_t)n * 8, &dd));
  PL_TRY(ws_get(WS_EIG_SEL, (size_t)k * n * 8, &dC));
  PL_TRY(ws_get(WS_DPP_VEC, (size_t)2 * (k > n ? k : n) * 8, &dcj));
  int nslices = (int)((k + 255) / 256);
  if (nslices > 16) nslices = 16;
  PL_TRY(ws_get(WS_DPP_TMP, (size_t)nslices * n * 8, &dtmp));
  PL_TRY(ws_get(WS_DPP_ITEMS, (size_t)k * 8, &ditems));
  CUDA_TRY(cudaEventRecord(g.ev[0], st));
  CUDA_TRY(cudaMemsetAsync(ditems, 0xff, (size_t)k * 8, st));
  CUDA_TRY(cudaEventRecord(g.ev[1], st));
  const unsigned nb = (unsigned)((n + 127) / 128);
  greedy_init_kernel<<<nb, 128, 0, st>>>(K, n, ld, (double*)dd);
  LAUNCH_CHECK();
  for (int64_t t = 0; t < k; ++t) {
    greedy_argmax_kernel<<<1, 1024, 0, st>>>((const double*)dd, n, t, (int64_t*)ditems);
    LAUNCH_CHECK();
    int sl = 1;
    if (t > 0) {
      greedy_gather_kernel<<<(unsigned)((t + 127) / 128), 128, 0, st>>>((const double*)dC, t, n, (const int64_t*)ditems, t, (double*)dcj);
      LAUNCH_CHECK();
      sl = (int)((t + 255) / 256);
      if (sl > nslices) sl = nslices;
      const int64_t sslice = (t + sl - 1) / sl;
      greedy_dot_partial_kernel<<<dim3(nb, (unsigned)sl), 128, 0, st>>>((const double*)dC, t, n, (const double*)dcj, sslice, (double*)dtmp);
      LAUNCH_CHECK();
    } else {
      CUDA_TRY(cudaMemsetAsync(dtmp, 0, (size_t)n * 8, st));
    }
    greedy_finish_kernel<<<nb, 128, 0, st>>>(K, n, ld, (const double*)dtmp, sl, (const int64_t*)ditems, t, (double*)dC + t * n, (double*)dd);
    LAUNCH_CHECK();
    greedy_retire_kernel<<<1, 32, 0, st>>>((const int64_t*)ditems, t, (double*)dd);
    LAUNCH_CHECK();
  }
  CUDA_TRY(cudaEventRecord(g.ev[2], st));
  CUDA_TRY(cudaMemcpyAsync(items, ditems, (size_t)k * 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(g.ev[3], st));
  return finish_timing();  // items[t] == -1 from the step at which the conditional variances are exhausted (numerical rank < k)
}


static int rmsd_host(const double* P, int64_t n, int64_t natoms, const double* box, double* D, int64_t ldd) {
  PL_TRY(need_ready());
  if (!P || !D || n < 0 || natoms <= 0 || ldd < n) return set_err(PL_EINVAL, "bad arguments");
  if (n > 65535) return set_err(PL_EINVAL, "at most 65535 configurations");
  if (box && !(box[0] != 0.0 && box[1] != 0.0 && box[2] != 0.0)) return set_err(PL_EINVAL, "box lengths must be non-zero");
  if (n == 0) return PL_OK;
  CUDA_TRY(cudaSetDevice(g.device));
  cudaStream_t st = g.stream;
  void *dP, *dPc = nullptr, *dD;
  const size_t pbytes = (size_t)n * natoms * 3 * 8;
  PL_TRY(ws_get(WS_POS, pbytes, &dP));
  PL_TRY(ws_get(WS_POS_D, (size_t)n * n * 8, &dD));
  CUDA_TRY(cudaEventRecord(g.ev[0], st));
  CUDA_TRY(cudaMemcpyAsync(dP, P, pbytes, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaEventRecord(g.ev[1], st));
  dim3 grid((unsigned)((n + 127) / 128), (unsigned)n);
  if (box) {
    rmsd_periodic_kernel<<<grid, 128, 0, st>>>((const double*)dP, n, natoms, box[0], box[1], box[2], (double*)dD, n);
    LAUNCH_CHECK();
  } else {
    PL_TRY(ws_get(WS_POS_C, pbytes, &dPc));
    centre_positions_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>((const double*)dP, n, natoms, (double*)dPc);
    LAUNCH_CHECK();
    rmsd_kabsch_kernel<<<grid, 128, 0, st>>>((const double*)dPc, n, natoms, (double*)dD, n);
    LAUNCH_CHECK();
  }
  CUDA_TRY(cudaEventRecord(g.ev[2], st));
  CUDA_TRY(cudaMemcpy2DAsync(D, (size_t)ldd * 8, dD, (size_t)n * 8, (size_t)n * 8, (size_t)n, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(g.ev[3], st));
  return finish_timing();
}

int pl_rmsd_periodic(const double* P, int64_t n, int64_t natoms, const double* box_lengths, double* D, int64_t ldd) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!box_lengths) return set_err(PL_EINVAL, "box_lengths is NULL");
  return rmsd_host(P, n, natoms, box_lengths, D, ldd);
}

int pl_rmsd_kabsch(const double* P, int64_t n, int64_t natoms, double* D, int64_t ldd) {
  std::lock_guard<std::mutex> lk(g_mu);
  return rmsd_host(P, n, natoms, nullptr, D, ldd);
}


int pl_metrics_dev(const double* dxp, const double* dx, int64_t n, double* dout, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  return metrics_dev_impl(dxp, dx, n, dout, (cudaStream_t)stream);
}

int pl_metrics(const double* x_pred, const double* x, int64_t n, double* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!x_pred || !x || !out || n <= 0) return set_err(PL_EINVAL, "bad arguments");
  CUDA_TRY(cudaSetDevice(g.device));
  CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
  double *dxp, *dx;
  PL_TRY(h2d_vector(x_pred, n, WS_H_IN2, &dxp));
  PL_TRY(h2d_vector(x, n, WS_H_IN3, &dx));
  void* dout;
  PL_TRY(ws_get(WS_H_OUT2, 64, &dout));
  CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
  PL_TRY(metrics_dev_impl(dxp, dx, n, (double*)dout, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
  CUDA_TRY(cudaMemcpyAsync(out, dout, 32, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[3], g.stream));
  return finish_timing();
}

}  // extern "C"

#include "multi.inl"  // NCCL (dlopen), pl_init_all / pl_comm_*, the sharded entry points and the resident Gram handle
