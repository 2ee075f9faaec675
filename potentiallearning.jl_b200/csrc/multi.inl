// multi.inl — several GPUs behind the C ABI (included at the end of plb200_api.cu; same translation unit).
//
// The reference has no distributed code at all; the path shards as section 8(e) says:
//   * normal equations / covariance (linear-learn.jl:253, dimension_reduction.jl:12): ROWS are sharded, every GPU runs the fused SYRK on its
//     shard, ONE all-reduce of the bucket [S ; v ; c ; s] (K^2 + 2K + 2 doubles) adds the partials, the finish kernel runs once;
//   * kernel matrices (kernels.jl:211-223): output tiles are independent — tile-column ranges (dense host result) or round-robin tiles
//     (resident, tile-packed: pl_gram_resident) per GPU, X replicated, no collective on the result.
// Two ways to drive several GPUs, both inside this library so that a Julia session needs no MPI and no Python:
//   * pl_init_all(ndev, ids): ONE process drives all GPUs. One context (stream, workspace, NCCL communicator from ncclCommInitAll) per
//     device; the host entry points (pl_normal_eq, pl_normal_eq_blocks, pl_cov, pl_gram_dot/rbf/imq) shard large inputs transparently,
//     one host thread per device, every device fed over its own PCIe link;
//   * pl_comm_init_rank(nranks, rank, id): one process per GPU (torchrun, mpirun); the *_sharded entry points take this rank's shard.
// NCCL is resolved with dlopen (like cuSOLVER): libplb200.so has no link-time dependency on it, and inside a process that already
// loaded a libnccl.so.2 (torch) the same copy is used.
#include <string>
#include <thread>

namespace {

typedef struct ncclComm* ncclComm_t;
struct ncclUniqueIdPl {
  char internal[128];  // NCCL_UNIQUE_ID_BYTES
};
constexpr int NCCL_SUCCESS = 0, NCCL_DOUBLE = 8, NCCL_SUM = 0;

struct Nccl {
  void* dl = nullptr;
  int version = 0;
  int (*GetVersion)(int*) = nullptr;
  int (*GetUniqueId)(ncclUniqueIdPl*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueIdPl, int) = nullptr;
  int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
} nc;

int nccl_ready() {
  if (nc.dl) return PL_OK;
  const char* env = getenv("PLB200_NCCL");
  const char* cands[] = {env, "libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
  void* dl = nullptr;
  for (const char* c : cands) {
    if (!c || !*c) continue;
    dl = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (dl) break;
  }
  if (!dl) return set_err(PL_ECUDA, "cannot load libnccl (%s); set PLB200_NCCL to its path", dlerror());
#define NC_SYM(field, name)                                                      \
  do {                                                                           \
    *(void**)(&nc.field) = dlsym(dl, name);                                      \
    if (!nc.field) return set_err(PL_ECUDA, "libnccl does not export %s", name); \
  } while (0)
  NC_SYM(GetVersion, "ncclGetVersion");
  NC_SYM(GetUniqueId, "ncclGetUniqueId");
  NC_SYM(CommInitRank, "ncclCommInitRank");
  NC_SYM(CommInitAll, "ncclCommInitAll");
  NC_SYM(CommDestroy, "ncclCommDestroy");
  NC_SYM(AllReduce, "ncclAllReduce");
  NC_SYM(AllGather, "ncclAllGather");
  NC_SYM(GroupStart, "ncclGroupStart");
  NC_SYM(GroupEnd, "ncclGroupEnd");
  NC_SYM(GetErrorString, "ncclGetErrorString");
#undef NC_SYM
  nc.GetVersion(&nc.version);
  nc.dl = dl;
  return PL_OK;
}

#define NCCL_TRY(expr)                                                                                          \
  do {                                                                                                          \
    int _r = (expr);                                                                                            \
    if (_r != NCCL_SUCCESS) return set_err(PL_ECUDA, "%s failed: %s (%s:%d)", #expr, nc.GetErrorString(_r), __FILE__, __LINE__); \
  } while (0)

// one process per GPU: the communicator of pl_comm_init_rank lives on context 0
int g_nranks = 1, g_rank = 0;
cudaEvent_t g_coll_ev[PL_MAX_DEV][2] = {};

// ---- the exchange step over PEER MEMORY (NVLink / NVSwitch loads and stores), fused behind the SYRK's own reduce kernel ---------------------
// Every rank owns a WINDOW [flags | bucket | result] that all other ranks map (cudaDeviceEnablePeerAccess inside one process,
// cudaIpc handles between processes). The second-stage kernel of the SYRK (syrk_reduce_kernel) writes its K x K sums straight into the
// bucket; then ONE kernel per rank does reduce-scatter + all-gather: after a flag barrier it loads its 1/n slice of every rank's bucket,
// adds them in rank order 0..n-1 and stores the sums into every rank's result buffer; a second flag barrier (peer_wait_kernel) releases
// the finish kernel, which reads the local result. 2 (n-1)/n x 8 K^2 bytes cross NVLink per rank (3.5 MB at K = 500, n = 8) instead of
// a ring's latency chain, and every element is summed by exactly one rank in a fixed order: all ranks hold IDENTICAL bits and the result
// is exactly symmetric. NCCL stays the fallback (no peer access, buckets above the window, PLB200_PEER_ALLREDUCE=0) and the transport
// of the large block fetches.
constexpr size_t PEER_CAP = (size_t)4 << 20;   // doubles per bucket / result region (32 MB: K up to ~2000)
constexpr size_t PEER_FLAG_BYTES = 4096;
struct PeerArgs {
  char* win[PL_MAX_DEV];  // window base of every rank as mapped in THIS process / device
  int n, me;
  unsigned long long epoch;
};
struct PeerWin {
  bool on = false, ipc = false;
  PeerArgs a = {};
};
PeerWin g_peer[PL_MAX_DEV];  // per context

__device__ __forceinline__ unsigned long long* peer_flag(char* win, int which, int rank) {
  return reinterpret_cast<unsigned long long*>(win) + which * PL_MAX_DEV + rank;
}
__device__ __forceinline__ double* peer_bucket(char* win) { return reinterpret_cast<double*>(win + PEER_FLAG_BYTES); }
__device__ __forceinline__ double* peer_result(char* win) { return reinterpret_cast<double*>(win + PEER_FLAG_BYTES) + PEER_CAP; }

// flag barrier `which` of this epoch: block 0 tells every rank "I am here" (release, system scope), every block waits until all
// ranks have said so in OUR window (acquire). A rank that never arrives (a crashed peer) traps after 30 s instead of hanging the GPU.
__device__ __forceinline__ void peer_barrier(const PeerArgs& a, int which) {
  if (blockIdx.x == 0 && threadIdx.x < a.n) {
    __threadfence_system();
    unsigned long long* f = peer_flag(a.win[threadIdx.x], which, a.me);
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(a.epoch) : "memory");
  }
  if (threadIdx.x < a.n) {
    const unsigned long long* f = peer_flag(a.win[a.me], which, threadIdx.x);
    unsigned long long v, t0 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v >= a.epoch) break;
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 30000000000ull) {
        printf("plb200: rank %d waited 30 s for rank %d at the peer-memory barrier (epoch %llu)\n", a.me, (int)threadIdx.x, a.epoch);
        __trap();
      }
    }
  }
  __syncthreads();
}

// reduce-scatter + all-gather of doubles [off, off + count) of the buckets (off even): see the block comment above
__global__ void peer_rsag_kernel(const PeerArgs a, size_t off, size_t count) {
  peer_barrier(a, 0);  // every rank's bucket is complete
  const size_t pairs = (count + 1) / 2, per = (pairs + a.n - 1) / a.n;
  const size_t lo = (size_t)a.me * per, hi = lo + per < pairs ? lo + per : pairs;
  for (size_t q = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < hi; q += (size_t)gridDim.x * blockDim.x) {
    double2 sum;
    {
      const double2* src = reinterpret_cast<const double2*>(peer_bucket(a.win[0]) + off) + q;
      asm volatile("ld.relaxed.sys.global.v2.f64 {%0,%1}, [%2];" : "=d"(sum.x), "=d"(sum.y) : "l"(src) : "memory");
    }
    for (int r = 1; r < a.n; ++r) {
      const double2* src = reinterpret_cast<const double2*>(peer_bucket(a.win[r]) + off) + q;
      double2 v;
      asm volatile("ld.relaxed.sys.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(src) : "memory");
      sum.x += v.x;
      sum.y += v.y;
    }
    for (int r = 0; r < a.n; ++r) reinterpret_cast<double2*>(peer_result(a.win[r]) + off)[q] = sum;
  }
  __threadfence_system();
}
// second barrier: every rank has stored its slices everywhere; what follows in the stream may read the local result
__global__ void peer_wait_kernel(const PeerArgs a) { peer_barrier(a, 1); }

// the bucket an operation should accumulate into: the peer window's when the exchange will run over peer memory, else workspace `slot`
int bucket_get(size_t count, int slot, double** out) {
  PeerWin& w = g_peer[g.index];
  if (w.on && count <= PEER_CAP) {
    *out = reinterpret_cast<double*>(w.a.win[w.a.me] + PEER_FLAG_BYTES);
    return PL_OK;
  }
  void* p;
  PL_TRY(ws_get(slot, count * 8, &p));
  *out = (double*)p;
  return PL_OK;
}

// Sum-all-reduce of `count` doubles over the ranks, enqueued on `st`. buf inside the peer window's bucket region: the peer-memory exchange,
// *out = the window's result region (same offset). Otherwise NCCL in place, *out = buf. Without a communicator: nothing, *out = buf.
int comm_allreduce(double* buf, size_t count, cudaStream_t st, double** out = nullptr) {
  if (out) *out = buf;
  if (!g.comm || count == 0) return PL_OK;
  cudaEvent_t* ev = g_coll_ev[g.index];
  if (!ev[0]) {
    CUDA_TRY(cudaEventCreate(&ev[0]));
    CUDA_TRY(cudaEventCreate(&ev[1]));
  }
  CUDA_TRY(cudaEventRecord(ev[0], st));
  PeerWin& w = g_peer[g.index];
  double* mine = w.on ? reinterpret_cast<double*>(w.a.win[w.a.me] + PEER_FLAG_BYTES) : nullptr;
  if (w.on && buf >= mine && buf + count <= mine + PEER_CAP && (((size_t)(buf - mine)) & 1) == 0) {
    const size_t off = (size_t)(buf - mine);
    ++w.a.epoch;
    const size_t per = ((count + 1) / 2 + w.a.n - 1) / w.a.n;
    int grid = (int)((per + 255) / 256);
    if (grid > 2 * g.num_sms) grid = 2 * g.num_sms;
    if (grid < 1) grid = 1;
    peer_rsag_kernel<<<grid, 256, 0, st>>>(w.a, off, count);
    LAUNCH_CHECK();
    peer_wait_kernel<<<1, 32, 0, st>>>(w.a);
    LAUNCH_CHECK();
    if (out) *out = mine + PEER_CAP + off;
    else  // in-place contract of the generic entry: bring the sums back
      CUDA_TRY(cudaMemcpyAsync(buf, mine + PEER_CAP + off, count * 8, cudaMemcpyDeviceToDevice, st));
  } else {
    NCCL_TRY(nc.AllReduce(buf, buf, count, NCCL_DOUBLE, NCCL_SUM, (ncclComm_t)g.comm, st));
  }
  CUDA_TRY(cudaEventRecord(ev[1], st));
  return PL_OK;
}

bool peer_wanted() {
  const char* e = getenv("PLB200_PEER_ALLREDUCE");
  return !(e && e[0] == '0');
}

// single process: windows on every device, every pair peer-enabled. Any failure leaves NCCL as the exchange (not an error).
void peer_setup_single(int ndev) {
  if (!peer_wanted() || ndev < 2) return;
  for (int i = 0; i < ndev; ++i)
    for (int j = 0; j < ndev; ++j) {
      if (i == j) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, g_ctx[i].device, g_ctx[j].device) != cudaSuccess || !can) {
        cudaGetLastError();
        return;
      }
    }
  char* win[PL_MAX_DEV] = {};
  bool ok = true;
  for (int i = 0; i < ndev && ok; ++i) {
    ok = cudaSetDevice(g_ctx[i].device) == cudaSuccess;
    for (int j = 0; j < ndev && ok; ++j) {
      if (i == j) continue;
      const cudaError_t e = cudaDeviceEnablePeerAccess(g_ctx[j].device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
      cudaGetLastError();
    }
    const size_t bytes = PEER_FLAG_BYTES + 2 * PEER_CAP * 8;
    if (ok) ok = cudaMalloc((void**)&win[i], bytes) == cudaSuccess && cudaMemset(win[i], 0, bytes) == cudaSuccess;
  }
  if (ok)
    for (int i = 0; i < ndev; ++i) ok = ok && cudaSetDevice(g_ctx[i].device) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    for (int i = 0; i < ndev; ++i)
      if (win[i]) {
        cudaSetDevice(g_ctx[i].device);
        cudaFree(win[i]);
      }
    return;
  }
  for (int i = 0; i < ndev; ++i) {
    PeerWin& w = g_peer[i];
    w = PeerWin();
    for (int j = 0; j < ndev; ++j) w.a.win[j] = win[j];
    w.a.n = ndev;
    w.a.me = i;
    w.on = true;
  }
}

// one process per GPU: the windows are exchanged as CUDA IPC handles through the communicator itself (collective: every rank calls it).
// All ranks agree on the outcome: the peer exchange is used only if EVERY rank mapped every window.
int peer_setup_ranks(int nranks, int rank, cudaStream_t st) {
  PeerWin& w = g_peer[0];
  w = PeerWin();
  if (nranks < 2 || nranks > PL_MAX_DEV) return PL_OK;
  const bool want = peer_wanted();
  const size_t bytes = PEER_FLAG_BYTES + 2 * PEER_CAP * 8;
  char* mine = nullptr;
  cudaIpcMemHandle_t hd;
  memset(&hd, 0, sizeof(hd));
  double okflag = 0.0;
  if (want && cudaMalloc((void**)&mine, bytes) == cudaSuccess && cudaMemset(mine, 0, bytes) == cudaSuccess && cudaIpcGetMemHandle(&hd, mine) == cudaSuccess)
    okflag = 1.0;
  cudaGetLastError();
  // gather the handles and count the ranks that got this far
  void* dbuf;
  PL_TRY(ws_get(WS_PADC, (size_t)(nranks + 1) * sizeof(hd) + 64, &dbuf));
  char* dh = (char*)dbuf;
  double* dok = (double*)(dh + (((size_t)(nranks + 1) * sizeof(hd) + 15) & ~(size_t)15));
  CUDA_TRY(cudaMemcpyAsync(dh + (size_t)nranks * sizeof(hd), &hd, sizeof(hd), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(dok, &okflag, 8, cudaMemcpyHostToDevice, st));
  NCCL_TRY(nc.AllGather(dh + (size_t)nranks * sizeof(hd), dh, sizeof(hd), 0 /* ncclChar */, (ncclComm_t)g.comm, st));
  NCCL_TRY(nc.AllReduce(dok, dok, 1, NCCL_DOUBLE, NCCL_SUM, (ncclComm_t)g.comm, st));
  std::vector<cudaIpcMemHandle_t> all(nranks);
  double nok = 0.0;
  CUDA_TRY(cudaMemcpyAsync(all.data(), dh, (size_t)nranks * sizeof(hd), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(&nok, dok, 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  double mapped = 0.0;
  if (nok == (double)nranks) {
    mapped = 1.0;
    w.a.win[rank] = mine;
    for (int r = 0; r < nranks; ++r) {
      if (r == rank) continue;
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        mapped = 0.0;
        break;
      }
      w.a.win[r] = (char*)p;
    }
  }
  CUDA_TRY(cudaMemcpyAsync(dok, &mapped, 8, cudaMemcpyHostToDevice, st));
  NCCL_TRY(nc.AllReduce(dok, dok, 1, NCCL_DOUBLE, NCCL_SUM, (ncclComm_t)g.comm, st));
  CUDA_TRY(cudaMemcpyAsync(&nok, dok, 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (nok == (double)nranks) {
    w.a.n = nranks;
    w.a.me = rank;
    w.ipc = true;
    w.on = true;
  } else {
    for (int r = 0; r < nranks; ++r)
      if (r != rank && w.a.win[r]) cudaIpcCloseMemHandle(w.a.win[r]);
    if (mine) cudaFree(mine);
    cudaGetLastError();
    w = PeerWin();
  }
  return PL_OK;
}

void peer_teardown() {
  for (int i = 0; i < PL_MAX_DEV; ++i) {
    PeerWin& w = g_peer[i];
    if (!w.on) continue;
    if (g_ctx[i].ready) cudaSetDevice(g_ctx[i].device);
    if (w.ipc) {
      for (int r = 0; r < w.a.n; ++r)
        if (r != w.a.me && w.a.win[r]) cudaIpcCloseMemHandle(w.a.win[r]);
    }
    if (w.a.win[w.a.me]) cudaFree(w.a.win[w.a.me]);
    cudaGetLastError();
    w = PeerWin();
  }
}

// out = a * x + b * y (K-length glue of the sharded covariance: column sums n_r z + rho, delta = m - z)
__global__ void axpby_kernel(double a, const double* __restrict__ x, double b, const double* __restrict__ y, int64_t count, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = a * x[i] + b * y[i];
}

// Runs fn(device index) once per context: contexts 1.. on their own host threads (current context and CUDA device set), context 0
// on the calling thread. The first failure (lowest device) is reported through the caller's pl_last_error.
template <class F>
int for_each_device(F&& fn) {
  const int n = g_ndev;
  std::vector<int> rc(n, PL_OK);
  std::vector<std::string> msg(n);
  std::vector<std::thread> th;
  th.reserve(n);
  for (int i = 1; i < n; ++i)
    th.emplace_back([&, i]() {
      tl_ctx = &g_ctx[i];
      if (cudaSetDevice(g_ctx[i].device) != cudaSuccess) {
        rc[i] = PL_ECUDA;
        msg[i] = "cudaSetDevice failed on a worker thread";
        return;
      }
      rc[i] = fn(i);
      if (rc[i] != PL_OK) msg[i] = g_err;
    });
  Ctx* saved = tl_ctx;
  tl_ctx = &g_ctx[0];
  cudaSetDevice(g_ctx[0].device);
  rc[0] = fn(0);
  if (rc[0] != PL_OK) msg[0] = g_err;
  for (auto& t : th) t.join();
  tl_ctx = saved;
  cudaSetDevice(saved->ready ? saved->device : g_ctx[0].device);
  for (int i = 0; i < n; ++i)
    if (rc[i] != PL_OK) return set_err(rc[i], "device %d: %s", g_ctx[i].device, msg[i].c_str());
  return PL_OK;
}

void multi_shutdown() {
  for (int i = 0; i < PL_MAX_DEV; ++i)
    if (g_ctx[i].ready) {
      cudaSetDevice(g_ctx[i].device);
      cudaDeviceSynchronize();
    }
  peer_teardown();
  for (int i = 0; i < PL_MAX_DEV; ++i) {
    if (g_ctx[i].comm && nc.CommDestroy) {
      cudaSetDevice(g_ctx[i].device);
      nc.CommDestroy((ncclComm_t)g_ctx[i].comm);
    }
    g_ctx[i].comm = nullptr;
    for (auto& ev : g_coll_ev[i]) {
      if (ev) cudaEventDestroy(ev);
      ev = nullptr;
    }
  }
  g_nranks = 1;
  g_rank = 0;
}

// balanced contiguous range of part i of n over [0, total)
inline void split_range(int64_t total, int i, int n, int64_t* lo, int64_t* hi) {
  const int64_t base = total / n, rem = total % n;
  *lo = i * base + (i < rem ? i : rem);
  *hi = *lo + base + (i < rem ? 1 : 0);
}

// the sub-list of row blocks covering logical rows [r_lo, r_hi) of the concatenation
void slice_blocks(const double* const* blocks, const int64_t* rows, const int64_t* lds, int64_t nblocks, int64_t r_lo, int64_t r_hi,
                  std::vector<const double*>* ob, std::vector<int64_t>* orows, std::vector<int64_t>* olds) {
  int64_t pos = 0;
  for (int64_t k = 0; k < nblocks && pos < r_hi; ++k) {
    const int64_t b0 = pos, b1 = pos + rows[k];
    pos = b1;
    const int64_t lo = r_lo > b0 ? r_lo : b0, hi = r_hi < b1 ? r_hi : b1;
    if (hi <= lo) continue;
    ob->push_back(blocks[k] + (lo - b0) * lds[k]);
    orows->push_back(hi - lo);
    olds->push_back(lds[k]);
  }
}

// ---- normal equations over several GPUs of one process (pl_normal_eq / pl_normal_eq_blocks after pl_init_all) ----
// Device i takes the i-th slice of the energy rows and the i-th slice of the force rows (its shard is laid out [energy ; force] like
// the whole), streams it through normal_eq_stream_partial over its own PCIe link, and the buckets are summed by one all-reduce.
int normal_eq_multi(const double* const* blocks, const int64_t* rows, const int64_t* lds, int64_t nblocks, int64_t R, int64_t Kf, const double* w,
                    int64_t ne, double w_e, double w_f, const double* b, int ic, double lambda, double* AtWA, double* AtWb, int64_t chunk_rows) {
  const bool want_vec = ic || b;
  const int n = g_ndev;
  if (chunk_rows <= 0) {
    chunk_rows = ((int64_t)256 << 20) / (Kf * 8);
    if (chunk_rows < 4096) chunk_rows = 4096;
  }
  chunk_rows = (chunk_rows + 15) / 16 * 16;
  return for_each_device([&](int i) -> int {
    int64_t e0, e1, f0, f1;
    split_range(ne, i, n, &e0, &e1);
    split_range(R - ne, i, n, &f0, &f1);
    std::vector<const double*> sb;
    std::vector<int64_t> srows, slds;
    slice_blocks(blocks, rows, lds, nblocks, e0, e1, &sb, &srows, &slds);
    slice_blocks(blocks, rows, lds, nblocks, ne + f0, ne + f1, &sb, &srows, &slds);
    sb.push_back(nullptr);  // sentinel: the cursor of normal_eq_stream_partial may look one block ahead
    srows.push_back(-1);
    slds.push_back(Kf);
    const int64_t ne_i = e1 - e0, R_i = ne_i + (f1 - f0);
    std::vector<double> wv, bv;
    if (w) {
      wv.assign(w + e0, w + e1);
      wv.insert(wv.end(), w + ne + f0, w + ne + f1);
    }
    if (b) {
      bv.assign(b + e0, b + e1);
      bv.insert(bv.end(), b + ne + f0, b + ne + f1);
    }
    double* bucket = nullptr;
    PL_TRY(normal_eq_stream_partial(sb.data(), srows.data(), slds.data(), (int64_t)sb.size() - 1, R_i, Kf, w ? wv.data() : nullptr, ne_i, w_e, w_f,
                                    b ? bv.data() : nullptr, want_vec, chunk_rows, &bucket));
    double* sums = nullptr;
    PL_TRY(comm_allreduce(bucket, (size_t)(Kf * Kf + (want_vec ? 2 * Kf + 2 : 0)), g.stream, &sums));
    if (i == 0) return normal_eq_bucket_finish(sums, Kf, want_vec, ic, lambda, AtWA, AtWb);
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    return PL_OK;
  });
}

// ---- covariance of one row shard + the exchange: the body shared by cov_multi (one thread per device) and pl_cov_sharded* ----
// dX: this context's rows (device). n_total: rows over all shards. dmean: K doubles on the device (given mean, or receives the
// global mean). dC receives (1/n_total) sum over ALL shards, exactly symmetric. Everything is enqueued on st.
int cov_shard_dev(const double* dX, int64_t n_local, int64_t n_total, int64_t Kf, int64_t ld, double* dmean, int mean_given, int center, double* dC,
                  cudaStream_t st) {
  // S (K x K) and, behind it, the K column sums share one bucket: both exchanges can run over the peer window
  const size_t soff = ((size_t)(Kf * Kf) + 1) & ~(size_t)1;
  double* dS;
  PL_TRY(bucket_get(soff + (size_t)Kf, WS_H_S, &dS));
  double* sums = dS;
  if (center && !mean_given) {
    // one pass over the rows: shift z = mean of a prefix of THIS shard, S_z and rho from the SYRK; the column sums n_r z + rho are
    // all-reduced for the global mean (pca.m = sum(d)/length(d), pca_state.jl:35), every shard re-centres about it, one all-reduce adds
    void *z, *rho, *delta;
    PL_TRY(ws_get(WS_COV_Z, (size_t)Kf * 8, &z));
    PL_TRY(ws_get(WS_COV_RHO, (size_t)Kf * 8, &rho));
    PL_TRY(ws_get(WS_COV_DELTA, (size_t)Kf * 8, &delta));
    double* csum = dS + soff;
    const unsigned kb = (unsigned)((Kf + 255) / 256);
    if (n_local > 0) {
      const int64_t prefix = n_local < 65536 ? n_local : 65536;
      PL_TRY(colsum(dX, prefix, Kf, ld, (double)prefix, (double*)z, st));
      PL_TRY(cov_shifted_dev_impl(dX, n_local, Kf, ld, (const double*)z, (double*)dS, (double*)rho, st));
    } else {
      CUDA_TRY(cudaMemsetAsync(z, 0, (size_t)Kf * 8, st));
      CUDA_TRY(cudaMemsetAsync(rho, 0, (size_t)Kf * 8, st));
      CUDA_TRY(cudaMemsetAsync(dS, 0, (size_t)Kf * Kf * 8, st));
    }
    axpby_kernel<<<kb, 256, 0, st>>>((double)n_local, (const double*)z, 1.0, (const double*)rho, Kf, (double*)csum);
    LAUNCH_CHECK();
    double* csum_all = nullptr;
    PL_TRY(comm_allreduce(csum, (size_t)Kf, st, &csum_all));
    scale_div_kernel<<<kb, 256, 0, st>>>((const double*)csum_all, (double)n_total, Kf, dmean);
    LAUNCH_CHECK();
    axpby_kernel<<<kb, 256, 0, st>>>(1.0, (const double*)dmean, -1.0, (const double*)z, Kf, (double*)delta);
    LAUNCH_CHECK();
    PL_TRY(cov_recentre_dev_impl((double*)dS, Kf, (const double*)rho, (const double*)delta, (double)n_local, st));
  } else {
    PL_TRY(syrk_dev_impl(dX, n_local, Kf, ld, nullptr, 0, 1.0, 1.0, nullptr, center ? dmean : nullptr, (double*)dS, nullptr, st));
  }
  PL_TRY(comm_allreduce(dS, (size_t)(Kf * Kf), st, &sums));
  int blocks = (int)((Kf * Kf + 255) / 256);
  if (blocks > g.num_sms * 8) blocks = g.num_sms * 8;
  // mean(di*di') = sum / n from the j >= i triangle: exactly symmetric whatever order the all-reduce summed in
  sym_scale_div_kernel<<<blocks, 256, 0, st>>>((const double*)sums, Kf, (double)n_total, dC);
  LAUNCH_CHECK();
  return PL_OK;
}

// host rows of this context's shard -> covariance (see cov_shard_dev); `root` copies C (and the mean) back and takes the timing
int cov_shard_host(const double* X, int64_t n_local, int64_t n_total, int64_t Kf, int64_t ldx, double* mean_inout, int mean_given, int center,
                   double* C, bool root) {
  CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
  double* dX;
  int64_t ld;
  PL_TRY(h2d_matrix(X, n_local, Kf, ldx, WS_H_IN1, &dX, &ld));
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
  void* dC;
  PL_TRY(ws_get(WS_H_OUT1, (size_t)Kf * Kf * 8, &dC));
  CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
  PL_TRY(cov_shard_dev(dX, n_local, n_total, Kf, ld, dmean, mean_given, center, (double*)dC, g.stream));
  CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
  if (root) {
    CUDA_TRY(cudaMemcpyAsync(C, dC, (size_t)Kf * Kf * 8, cudaMemcpyDeviceToHost, g.stream));
    if (center && !mean_given) CUDA_TRY(cudaMemcpyAsync(mean_inout, dmean, (size_t)Kf * 8, cudaMemcpyDeviceToHost, g.stream));
  }
  CUDA_TRY(cudaEventRecord(g.ev[3], g.stream));
  return finish_timing();
}

int cov_multi(const double* X, int64_t n, int64_t Kf, int64_t ldx, double* mean_inout, int mean_given, int center, double* C) {
  const int nd = g_ndev;
  return for_each_device([&](int i) -> int {
    int64_t lo, hi;
    split_range(n, i, nd, &lo, &hi);
    return cov_shard_host(X + lo * ldx, hi - lo, n, Kf, ldx, mean_inout, mean_given, center, C, i == 0);
  });
}

// ---- symmetric kernel matrix, dense Symmetric(:U) host result, over several GPUs of one process ----
// Every device receives X over its own PCIe link, computes one contiguous share of the tile columns (equal work) and streams its
// finished row blocks straight into the caller's matrix: no collective, and the D2H copies (the end-to-end bottleneck) run in parallel.
int gram_sym_multi(int kernel_id, const double* X, int64_t n, int64_t d, int64_t ldx, int alpha, int cinv_kind, const double* cinv, double ell,
                   double beta, double* K, int64_t ldk) {
  const int nd = g_ndev;
  const int per = 4;  // column ranges per device (copy of range c overlaps the compute of range c + 1)
  const bool dist_kernel = (kernel_id == PL_KERNEL_RBF || kernel_id == PL_KERNEL_IMQ);
  if (dist_kernel && cinv_kind != PL_CINV_IDENTITY && !cinv) return set_err(PL_EINVAL, "cinv is NULL");
  return for_each_device([&](int i) -> int {
    CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
    double *dX, *dc = nullptr;
    int64_t ld;
    PL_TRY(h2d_matrix(X, n, d, ldx, WS_H_IN1, &dX, &ld));
    if (dist_kernel && cinv_kind == PL_CINV_DIAGONAL) PL_TRY(h2d_vector(cinv, d, WS_H_IN3, &dc));
    else if (dist_kernel && cinv_kind == PL_CINV_FULL) PL_TRY(h2d_vector(cinv, d * d, WS_H_IN3, &dc));
    return gram_sym_stream(kernel_id, dX, n, ld, d, alpha, cinv_kind, dc, ell, beta, K, ldk, i * per, (i + 1) * per, nd * per);
  });
}

// ---- resident, tile-packed kernel matrix (config 5: N = 200 000 -> 160 GB of upper triangle, 8 x 20 GB shards that never leave the GPUs) ----
struct GramResident {
  int64_t n = 0;
  int T = 0;
  int nparts = 1;            // shards: devices of this process (pl_init_all) or ranks (pl_comm_init_rank)
  int nlocal = 0;            // shards held by this process
  int part[PL_MAX_DEV] = {};  // global shard id of local shard i (= context i)
  double* tiles[PL_MAX_DEV] = {};
  int64_t ntiles[PL_MAX_DEV] = {};
};

// inverse of gram_slot_to_tile for the symmetric (folded) enumeration: slot of tile (ti <= tj)
__host__ __device__ __forceinline__ int64_t gram_tile_to_slot(int T, int ti, int tj) {
  if (2 * ti <= T - 1) return (int64_t)ti * (T + 1) + (tj - ti);
  const int p = T - 1 - ti;
  return (int64_t)p * (T + 1) + (T - p) + (tj - ti);
}

// stage[(i - i0) * lds + (j - j0)] = K[i][j] for the elements of block [i0, i1) x [j0, j1) whose tile this shard owns (mirrored for j < i)
__global__ void gram_fetch_kernel(const double* __restrict__ tiles, int T, int part, int nparts, int64_t i0, int64_t i1, int64_t j0, int64_t j1,
                                  double* __restrict__ stage, int64_t lds) {
  const int64_t w = j1 - j0, total = (i1 - i0) * w;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i0 + e / w, j = j0 + e % w;
    const int64_t a = i <= j ? i : j, b = i <= j ? j : i;  // stored element [a][b], a <= b
    const int ti = (int)(a / GRAM_BM), tj = (int)(b / GRAM_BN);
    const int64_t slot = gram_tile_to_slot(T, ti, tj);
    if (slot % nparts != part) continue;
    stage[(i - i0) * lds + (j - j0)] = tiles[(slot / nparts) * (int64_t)(GRAM_BM * GRAM_BN) + (a % GRAM_BM) * GRAM_BN + (b % GRAM_BN)];
  }
}

}  // namespace

extern "C" {

int pl_init_all(int ndev, const int* device_ids) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_ctx[0].ready) {
    if (g_ndev == ndev || ndev <= 0) return PL_OK;
    return set_err(PL_EINVAL, "plb200 is already initialised with %d device(s); call pl_finalize first", g_ndev);
  }
  int count = 0;
  PL_TRY(device_count_checked(&count));
  if (ndev <= 0) ndev = count;
  if (ndev > count || ndev > PL_MAX_DEV) return set_err(PL_EINVAL, "%d devices requested, %d available (at most %d)", ndev, count, PL_MAX_DEV);
  int ids[PL_MAX_DEV];
  for (int i = 0; i < ndev; ++i) {
    ids[i] = device_ids ? device_ids[i] : i;
    if (ids[i] < 0 || ids[i] >= count) return set_err(PL_EINVAL, "device %d out of range (%d devices)", ids[i], count);
    for (int j = 0; j < i; ++j)
      if (ids[j] == ids[i]) return set_err(PL_EINVAL, "device %d listed twice", ids[i]);
  }
  for (int i = 0; i < ndev; ++i) {
    const int rc = ctx_create(i, ids[i]);
    if (rc != PL_OK) {
      for (int j = i; j >= 0; --j) ctx_destroy(j);
      return rc;
    }
  }
  g_ndev = ndev;
  tl_ctx = &g_ctx[0];
  if (ndev > 1) {
    ncclComm_t comms[PL_MAX_DEV];
    int rc = nccl_ready();
    if (rc == PL_OK) {
      const int r = nc.CommInitAll(comms, ndev, ids);
      if (r != NCCL_SUCCESS) rc = set_err(PL_ECUDA, "ncclCommInitAll over %d devices failed: %s", ndev, nc.GetErrorString(r));
    }
    if (rc != PL_OK) {
      for (int j = ndev - 1; j >= 0; --j) ctx_destroy(j);
      g_ndev = 0;
      return rc;
    }
    for (int i = 0; i < ndev; ++i) g_ctx[i].comm = comms[i];
    peer_setup_single(ndev);  // peer windows for the exchange over NVLink; NCCL remains the fallback
    // pinned staging rings up front (see HostStager::start): 16 host threads over all devices
    const int per = 16 / ndev < 2 ? 2 : 16 / ndev;
    for (int i = 0; i < ndev; ++i) {
      CUDA_TRY(cudaSetDevice(g_ctx[i].device));
      CUDA_TRY(g_ctx[i].stager.start(per));
    }
  }
  CUDA_TRY(cudaSetDevice(g_ctx[0].device));
  return PL_OK;
}

int pl_device_count(void) {
  return g_ndev;
}

int pl_comm_unique_id(void* out128) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!out128) return set_err(PL_EINVAL, "null pointer");
  PL_TRY(nccl_ready());
  ncclUniqueIdPl id;
  NCCL_TRY(nc.GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return PL_OK;
}

int pl_comm_init_rank(int nranks, int rank, const void* id128) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (g_ndev != 1) return set_err(PL_EINVAL, "pl_comm_init_rank is for one process per GPU (pl_init); this process drives %d devices", g_ndev);
  if (nranks < 1 || rank < 0 || rank >= nranks || !id128) return set_err(PL_EINVAL, "bad arguments");
  if (g.comm) return set_err(PL_EINVAL, "a communicator already exists: call pl_comm_destroy first");
  PL_TRY(nccl_ready());
  ncclUniqueIdPl id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  NCCL_TRY(nc.CommInitRank(&c, nranks, id, rank));
  g.comm = c;
  g_nranks = nranks;
  g_rank = rank;
  return peer_setup_ranks(nranks, rank, g.stream);  // collective: IPC-mapped peer windows (falls back to NCCL on every rank alike)
}

int pl_comm_peer_enabled(void) {
  return g_peer[0].on ? 1 : 0;
}

int pl_comm_info(int* nranks, int* rank, int* nccl_version) {
  if (nranks) *nranks = g_ctx[0].comm ? (g_ndev > 1 ? g_ndev : g_nranks) : 1;
  if (rank) *rank = g_rank;
  if (nccl_version) *nccl_version = nc.version;
  return PL_OK;
}

int pl_comm_destroy(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_ndev > 1) return set_err(PL_EINVAL, "the communicators of pl_init_all live until pl_finalize");
  if (g_ctx[0].comm) {
    CUDA_TRY(cudaSetDevice(g_ctx[0].device));
    CUDA_TRY(cudaDeviceSynchronize());
    peer_teardown();
    nc.CommDestroy((ncclComm_t)g_ctx[0].comm);
    g_ctx[0].comm = nullptr;
  }
  g_nranks = 1;
  g_rank = 0;
  return PL_OK;
}

int pl_last_coll_ms(double* ms) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  cudaEvent_t* ev = g_coll_ev[g.index];
  if (!ev[0]) return set_err(PL_EINVAL, "no collective has run on this device");
  CUDA_TRY(cudaEventSynchronize(ev[1]));
  float t = 0;
  CUDA_TRY(cudaEventElapsedTime(&t, ev[0], ev[1]));
  if (ms) *ms = t;
  return PL_OK;
}

int pl_allreduce_dev(double* dbuf, int64_t count, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (count < 0 || (count > 0 && !dbuf)) return set_err(PL_EINVAL, "bad arguments");
  return comm_allreduce(dbuf, (size_t)count, (cudaStream_t)stream);
}

int pl_normal_eq_sharded_dev(const double* dA, int64_t R_local, int64_t Kf, int64_t lda, const double* dw, int64_t n_energy_rows_local, double w_e,
                             double w_f, const double* db, int add_intercept, double lambda, double* dAtWA, double* dAtWb, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!dA || !dAtWA || R_local < 0 || Kf <= 0 || lda < Kf) return set_err(PL_EINVAL, "bad arguments");
  if (dAtWb && !db) return set_err(PL_EINVAL, "AtWb requested without b");
  const int ic = add_intercept ? 1 : 0;
  const bool want_vec = ic || db;
  cudaStream_t st = (cudaStream_t)stream;
  // the SYRK's second-stage kernel writes its sums straight into the exchange bucket (the peer window when it is mapped)
  double* dS;
  PL_TRY(bucket_get((size_t)(Kf * Kf + 2 * Kf + 2), WS_NE_SACC, &dS));
  double* dvec = want_vec ? dS + Kf * Kf : nullptr;
  PL_TRY(syrk_dev_impl(dA, R_local, Kf, lda, dw, n_energy_rows_local, w_e, w_f, db, nullptr, dS, dvec, st));
  double* sums = nullptr;
  PL_TRY(comm_allreduce(dS, (size_t)(Kf * Kf + (want_vec ? 2 * Kf + 2 : 0)), st, &sums));
  return finish_dev_impl(sums, want_vec ? sums + Kf * Kf : nullptr, Kf, ic, lambda, dAtWA, dAtWb, st);
}

int pl_cov_sharded_dev(const double* dX, int64_t n_local, int64_t n_total, int64_t Kf, int64_t ldx, double* dmean_inout, int mean_given, int center,
                       double* dC, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!dX || !dC || n_local < 0 || n_total <= 0 || Kf <= 0 || ldx < Kf) return set_err(PL_EINVAL, "bad arguments");
  if (center && !dmean_inout) return set_err(PL_EINVAL, "centred covariance needs dmean_inout");
  cudaStream_t st = (cudaStream_t)stream;
  const double* X;
  int64_t ld;
  PL_TRY(tma_operand(dX, n_local, Kf, ldx, WS_PAD1, st, &X, &ld));
  return cov_shard_dev(X, n_local, n_total, Kf, ld, dmean_inout, mean_given, center, dC, st);
}

int pl_normal_eq_sharded(const double* A_local, int64_t R_local, int64_t Kf, int64_t lda, const double* w, int64_t n_energy_rows_local, double w_e,
                         double w_f, const double* b, int add_intercept, double lambda, double* AtWA, double* AtWb) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (g_ndev != 1) return set_err(PL_EINVAL, "pl_normal_eq_sharded is for one process per GPU; after pl_init_all call pl_normal_eq with all the rows");
  if (!AtWA || R_local < 0 || Kf <= 0 || lda < Kf || (R_local > 0 && !A_local)) return set_err(PL_EINVAL, "bad arguments");
  if (AtWb && !b) return set_err(PL_EINVAL, "AtWb requested without b");
  const int ic = add_intercept ? 1 : 0;
  const bool want_vec = ic || b;
  int64_t chunk_rows = (((int64_t)512 << 20) / (Kf * 8) + 15) / 16 * 16;
  const double* blk[2] = {A_local, nullptr};
  const int64_t rows[2] = {R_local, -1}, lds[2] = {lda, Kf};
  int64_t ne = n_energy_rows_local < 0 ? 0 : (n_energy_rows_local > R_local ? R_local : n_energy_rows_local);
  double* bucket = nullptr;
  PL_TRY(normal_eq_stream_partial(blk, rows, lds, 1, R_local, Kf, w, ne, w_e, w_f, b, want_vec, chunk_rows, &bucket));
  double* sums = nullptr;
  PL_TRY(comm_allreduce(bucket, (size_t)(Kf * Kf + (want_vec ? 2 * Kf + 2 : 0)), g.stream, &sums));
  return normal_eq_bucket_finish(sums, Kf, want_vec, ic, lambda, AtWA, AtWb);
}

int pl_cov_sharded(const double* X_local, int64_t n_local, int64_t n_total, int64_t Kf, int64_t ldx, double* mean_inout, int mean_given, int center,
                   double* C) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (g_ndev != 1) return set_err(PL_EINVAL, "pl_cov_sharded is for one process per GPU; after pl_init_all call pl_cov with all the rows");
  if (!C || n_local < 0 || n_total <= 0 || Kf <= 0 || ldx < Kf || (n_local > 0 && !X_local)) return set_err(PL_EINVAL, "bad arguments");
  if (center && !mean_inout) return set_err(PL_EINVAL, "centred covariance needs mean_inout");
  return cov_shard_host(X_local, n_local, n_total, Kf, ldx, mean_inout, mean_given, center, C, true);
}

// ---- resident kernel matrix -------------------------------------------------------------------------------------------------------
int pl_gram_resident(int kernel_id, const double* X, int64_t n, int64_t d, int64_t ldx, int alpha, int cinv_kind, const double* cinv, double ell,
                     double beta, void** handle) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!X || !handle || n <= 0 || d <= 0 || ldx < d) return set_err(PL_EINVAL, "bad arguments");
  const bool dist_kernel = (kernel_id == PL_KERNEL_RBF || kernel_id == PL_KERNEL_IMQ);
  if (dist_kernel && cinv_kind != PL_CINV_IDENTITY && !cinv) return set_err(PL_EINVAL, "cinv is NULL");
  GramResident* H = new GramResident();
  H->n = n;
  H->T = (int)((n + GRAM_BM - 1) / GRAM_BM);
  const bool ranks = g_ndev == 1 && g_ctx[0].comm;
  H->nparts = ranks ? g_nranks : g_ndev;
  H->nlocal = ranks ? 1 : g_ndev;
  const int64_t slots = gram_num_slots(H->T, H->T, 1);
  int rc = for_each_device([&](int i) -> int {
    const int part = ranks ? g_rank : i;
    H->part[i] = part;
    const int64_t mine = (slots - part + H->nparts - 1) / H->nparts;
    H->ntiles[i] = mine > 0 ? mine : 0;
    void* buf = nullptr;
    cudaError_t e = cudaMalloc(&buf, (size_t)(mine > 0 ? mine : 1) * GRAM_BM * GRAM_BN * 8);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return set_err(PL_ENOMEM, "resident kernel matrix: %lld tiles (%.1f GB) do not fit on device %d", (long long)mine, mine * 131072.0 / 1e9, g.device);
    }
    H->tiles[i] = (double*)buf;
    CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
    double *dX, *dc = nullptr;
    int64_t ld;
    PL_TRY(h2d_matrix(X, n, d, ldx, WS_H_IN1, &dX, &ld));
    if (dist_kernel && cinv_kind == PL_CINV_DIAGONAL) PL_TRY(h2d_vector(cinv, d, WS_H_IN3, &dc));
    else if (dist_kernel && cinv_kind == PL_CINV_FULL) PL_TRY(h2d_vector(cinv, d * d, WS_H_IN3, &dc));
    CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
    PL_TRY(gram_dev_impl(kernel_id, dX, n, ld, nullptr, 0, 0, d, alpha, cinv_kind, dc, ell, beta, H->tiles[i], GRAM_BN, PL_FILL_FULL, part, H->nparts, 1, g.stream));
    CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
    CUDA_TRY(cudaEventRecord(g.ev[3], g.stream));
    return finish_timing();
  });
  if (rc != PL_OK) {
    for (int i = 0; i < H->nlocal; ++i)
      if (H->tiles[i]) {
        cudaSetDevice(g_ctx[i].device);
        cudaFree(H->tiles[i]);
      }
    cudaSetDevice(g_ctx[0].device);
    delete H;
    return rc;
  }
  *handle = H;
  return PL_OK;
}

int pl_gram_resident_info(void* handle, int64_t* n, int* nparts, int64_t* tiles_local, int64_t* bytes_local) {
  if (!handle) return set_err(PL_EINVAL, "null handle");
  const GramResident* H = (const GramResident*)handle;
  int64_t t = 0;
  for (int i = 0; i < H->nlocal; ++i) t += H->ntiles[i];
  if (n) *n = H->n;
  if (nparts) *nparts = H->nparts;
  if (tiles_local) *tiles_local = t;
  if (bytes_local) *bytes_local = t * (int64_t)(GRAM_BM * GRAM_BN * 8);
  return PL_OK;
}

// device pointer of the tile-packed shard held by local device `local_index` (for callers that keep working on the device)
int pl_gram_resident_shard(void* handle, int local_index, int* part, int64_t* ntiles, void** dtiles) {
  if (!handle) return set_err(PL_EINVAL, "null handle");
  const GramResident* H = (const GramResident*)handle;
  if (local_index < 0 || local_index >= H->nlocal) return set_err(PL_EINVAL, "shard %d out of range (%d local shards)", local_index, H->nlocal);
  if (part) *part = H->part[local_index];
  if (ntiles) *ntiles = H->ntiles[local_index];
  if (dtiles) *dtiles = H->tiles[local_index];
  return PL_OK;
}

int pl_gram_fetch_block(void* handle, int64_t i0, int64_t i1, int64_t j0, int64_t j1, double* out, int64_t ldo) {
  std::lock_guard<std::mutex> lk(g_mu);
  PL_TRY(need_ready());
  if (!handle || !out) return set_err(PL_EINVAL, "null pointer");
  const GramResident* H = (const GramResident*)handle;
  if (i0 < 0 || j0 < 0 || i1 > H->n || j1 > H->n || i1 < i0 || j1 < j0 || ldo < j1 - j0) return set_err(PL_EINVAL, "block out of range");
  const int64_t h = i1 - i0, w = j1 - j0;
  if (h == 0 || w == 0) return PL_OK;
  if ((double)h * (double)w > 268435456.0) return set_err(PL_EINVAL, "block larger than 2 GB: fetch it in pieces");
  // every shard writes the elements it owns into a zeroed staging block; one all-reduce (x + 0 + ... + 0 is exact) assembles it
  return for_each_device([&](int i) -> int {
    void* stage;
    PL_TRY(ws_get(WS_H_OUT1, (size_t)h * w * 8, &stage));
    CUDA_TRY(cudaEventRecord(g.ev[0], g.stream));
    CUDA_TRY(cudaEventRecord(g.ev[1], g.stream));
    CUDA_TRY(cudaMemsetAsync(stage, 0, (size_t)h * w * 8, g.stream));
    int64_t blocks = (h * w + 255) / 256;
    if (blocks > (int64_t)g.num_sms * 16) blocks = (int64_t)g.num_sms * 16;
    gram_fetch_kernel<<<(unsigned)blocks, 256, 0, g.stream>>>(H->tiles[i], H->T, H->part[i], H->nparts, i0, i1, j0, j1, (double*)stage, w);
    LAUNCH_CHECK();
    if (H->nparts > 1) PL_TRY(comm_allreduce((double*)stage, (size_t)(h * w), g.stream));
    CUDA_TRY(cudaEventRecord(g.ev[2], g.stream));
    if (i == 0) CUDA_TRY(cudaMemcpy2DAsync(out, (size_t)ldo * 8, stage, (size_t)w * 8, (size_t)w * 8, (size_t)h, cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaEventRecord(g.ev[3], g.stream));
    return finish_timing();
  });
}

int pl_gram_free(void* handle) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!handle) return PL_OK;
  GramResident* H = (GramResident*)handle;
  for (int i = 0; i < H->nlocal; ++i)
    if (H->tiles[i] && g_ctx[i].ready) {
      cudaSetDevice(g_ctx[i].device);
      cudaDeviceSynchronize();
      cudaFree(H->tiles[i]);
    }
  if (g_ctx[0].ready) cudaSetDevice(g_ctx[0].device);
  delete H;
  return PL_OK;
}

}  // extern "C"
