/*
 * plb200.h — C ABI of libplb200.so: the B200 (sm_100a) descriptor linear-algebra hot path of
 * PotentialLearning.jl (reference v0.2.7, citations are relative to /root/reference).
 *
 * This is the drop-in boundary. The reference has no FFI of its own (it is pure Julia and the
 * boundary is multiple dispatch on exported generics), so every entry point below names the
 * reference method body it replaces; INTEGRATION.md shows the Julia `ccall` stub per entry.
 *
 * Conventions
 *  - all arithmetic is IEEE fp64; every function returns 0 on success, a PL_E* code otherwise;
 *    the message of the last failure on the calling thread is pl_last_error().
 *  - "host" entry points take HOST pointers owned by the caller. They copy in, launch, copy out and
 *    are synchronous; no host pointer is retained after return (Julia only roots arrays for the
 *    duration of a ccall).
 *  - "_dev" entry points take DEVICE pointers (16-byte aligned, leading dimension even) plus a
 *    cudaStream_t passed as void*; they only enqueue work on that stream (no sync) and are what
 *    the multi-GPU host code and the benchmarks use. They run on the library's device whatever device
 *    the caller has current, and they SHARE the library's scratch buffers (SYRK partials, schedule table,
 *    centred copies, row statistics): enqueue them on ONE stream at a time — calls on two streams must be
 *    ordered by the caller (event / synchronise), also across different entry points.
 *  - Matrices of descriptors are ROW-MAJOR n x d with leading dimension ld (elements). This is
 *    byte-identical to Julia's column-major d x n produced by reduce(hcat, F).
 *  - There is NO CPU fallback: without an sm_100 device pl_init fails and every entry errors.
 */
#ifndef PLB200_H
#define PLB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes */
#define PL_OK 0
#define PL_EINVAL 1   /* bad argument (shape, stride, alignment, enum) */
#define PL_ECUDA 2    /* CUDA runtime / driver error (message has the CUDA string) */
#define PL_ENODEV 3   /* no sm_100 device / library not initialised */
#define PL_ENOMEM 4   /* device allocation failed */

/* kernel ids (src/Kernels/kernels.jl) */
#define PL_KERNEL_LINEAR 0 /* plain Gram X1*X2' (no epilogue); building block, e.g. X*Cinv      */
#define PL_KERNEL_DOT 1    /* DotProduct: (a.b / sqrt((a.a)(b.b)))^alpha      kernels.jl:21-36  */
#define PL_KERNEL_RBF 2    /* RBF(Euclidean): beta*exp(-0.5*d2/ell^2)         kernels.jl:51-75  */
#define PL_KERNEL_IMQ 3    /* InverseMultiquadric(Euclidean): (c2 + d2/ell^2)^(-1/2)  kernels.jl:139-164 ("next" row f4:
                              same Gram engine, one more epilogue). In the generic entries `beta` carries c2.           */

/* Euclidean Cinv kinds (src/Kernels/distances.jl:39-51) */
#define PL_CINV_IDENTITY 0 /* Euclidean(dim): 1.0*I(dim)                      distances.jl:43-45 */
#define PL_CINV_DIAGONAL 1 /* Diagonal{T}: cinv = d doubles                                      */
#define PL_CINV_FULL 2     /* Symmetric{T}: cinv = d x d (symmetric, so row/col-major agree)     */

/* which triangle(s) of a symmetric result are written. Bits refer to the ROW-MAJOR view out[i*ld+j]:
 * PL_FILL_RM_UPPER writes j>=i, PL_FILL_RM_LOWER writes j<=i. Julia's Symmetric(K) (uplo=:U, column-major,
 * kernels.jl:222) is PL_FILL_RM_LOWER of the same bytes; the unused triangle of the storage is written as 0,
 * exactly as the reference leaves it (zeros(n,n), kernels.jl:216). */
#define PL_FILL_RM_UPPER 1
#define PL_FILL_RM_LOWER 2
#define PL_FILL_FULL 3
#define PL_FILL_JULIA_U PL_FILL_RM_LOWER

/* ---- lifetime ------------------------------------------------------------------------------------------ */
/* Bind the library to one CUDA device (-1: the current device): one process per GPU. Multi-GPU jobs of that kind create a
 * communicator with pl_comm_init_rank and use the *_sharded entry points. Fails with PL_ENODEV unless the device is compute
 * capability 10.x. Idempotent. */
int pl_init(int device);
/* ONE process drives ndev GPUs (what a Julia session does): a context (streams, workspace) per device and one NCCL
 * communicator per device from ncclCommInitAll inside the library. device_ids == NULL: devices 0..ndev-1; ndev <= 0: every visible
 * device. Afterwards the HOST entry points shard large inputs over the devices transparently — pl_normal_eq / pl_normal_eq_blocks /
 * pl_cov shard rows (each device streams its shard over its own PCIe link, one all-reduce of K^2+2K+2 doubles adds the partials:
 * learn! linear-learn.jl:253, compute_eigen dimension_reduction.jl:12), pl_gram_dot / pl_gram_rbf / pl_gram_imq with PL_FILL_JULIA_U
 * shard tile columns (no collective, parallel D2H: KernelMatrix kernels.jl:211-223) — results identical in meaning and layout to one
 * device. The _dev entry points and everything else run on device_ids[0]. */
int pl_init_all(int ndev, const int* device_ids);
/* devices driven by this process (0 before pl_init / pl_init_all) */
int pl_device_count(void);
void pl_finalize(void);

/* ---- one process per GPU: the communicator (NCCL, loaded with dlopen; PLB200_NCCL overrides the library path) --------------- */
/* rank 0 calls pl_comm_unique_id and distributes the 128 bytes by any means (torch.distributed store, MPI, a file); then EVERY rank
 * calls pl_comm_init_rank (collective). nranks / rank as in ncclCommInitRank. */
int pl_comm_unique_id(void* out128);
int pl_comm_init_rank(int nranks, int rank, const void* id128);
int pl_comm_info(int* nranks, int* rank, int* nccl_version);
/* 1 when the K x K exchange of the sharded entry points runs over PEER MEMORY (every rank maps every other rank's exchange window:
 * cudaIpc between processes, peer access inside one process): the SYRK's second-stage kernel writes into the window, one kernel does
 * reduce-scatter + all-gather with NVLink loads / stores between two flag barriers, every rank ends with identical bits. 0: NCCL
 * (no peer access, bucket above the 32 MB window, or PLB200_PEER_ALLREDUCE=0). */
int pl_comm_peer_enabled(void);
int pl_comm_destroy(void);
/* milliseconds of the last all-reduce enqueued by the library on this process's first device (CUDA events around it) */
int pl_last_coll_ms(double* ms);
const char* pl_last_error(void);
/* milliseconds spent by the last host entry point: kernels (CUDA events), H2D and D2H copies. */
int pl_last_timing(double* kernel_ms, double* h2d_ms, double* d2h_ms);
/* wall-clock milliseconds the last host entry point spent inside the library, entry to return (>= the three above: the rest is host work) */
int pl_last_host_ms(double* ms);
/* number of plb200 CUDA kernels launched by this process so far (benchmarks report it as gpu_launches). */
int64_t pl_launch_count(void);
/* bytes of device workspace currently held by the library. */
int64_t pl_workspace_bytes(void);
/* Frees the library's device scratch (it is grown lazily and otherwise kept until pl_finalize). keep_resident != 0 keeps the resident
 * L-ensemble (pl_ell_*) and an open normal-equation accumulator (pl_neq_*); 0 drops them as well. */
int pl_workspace_trim(int keep_resident);
/* on != 0: the dominant DMMA kernel (Gram or SYRK) of every following operation is bracketed by CUDA events on the
 * stream it is launched on; pl_last_main_kernel_ms waits for it and returns its duration (roofline measurements). */
int pl_set_profiling(int on);
int pl_last_main_kernel_ms(double* ms);

/* Pinned, zero-filled host memory for result arrays (the shim wraps it as a Julia Array with a finalizer). The host entry points accept
 * any host pointer (pageable memory is staged through the library's own pinned ring and host threads), but only pinned memory is copied at
 * the PCIe rate: config 2 end to end takes 37 ms into a pinned K, 45 ms into pageable memory, ~100 ms into freshly calloc'ed pages.
 * NULL on failure (pl_last_error). */
void* pl_host_alloc(int64_t bytes);
int pl_host_free(void* p);

/* ---- Gram / kernel matrices: replaces KernelMatrix(F,k) kernels.jl:211-223 and (F1,F2,k) :229-244 ------ */

/* Symmetric DotProduct kernel matrix of n features of length d.
 *   K[i,j] = (x_i.x_j / sqrt((x_i.x_i)(x_j.x_j)))^alpha        (kernels.jl:35; alpha is the Int power)
 * K: n x n, leading dimension ldk, triangles per `fill`. */
int pl_gram_dot(const double* X, int64_t n, int64_t d, int64_t ldx, int alpha, double* K, int64_t ldk, int fill);

/* Symmetric RBF(Euclidean) kernel matrix.
 *   d2 = (x_i-x_j)' Cinv (x_i-x_j)  (distances.jl:58-60);  K[i,j] = beta*exp(-0.5*d2/ell^2) (kernels.jl:73-74).
 * RBF.alpha is NOT applied (the reference never applies it); K[i,i] == beta exactly.
 * Accuracy: <= 1e-11 relative against the reference formula wherever K > 2.3e-308; entries the reference evaluates as denormal or 0
 * (-0.5*d2/ell^2 + ln(beta) < -708.4) come out as some number <= 4.5e-308. A NaN feature vector makes its own row and column NaN and
 * nothing else (as in the reference).
 * cinv: NULL for PL_CINV_IDENTITY, d doubles for DIAGONAL, d*d doubles for FULL. */
int pl_gram_rbf(const double* X, int64_t n, int64_t d, int64_t ldx, int cinv_kind, const double* cinv, double ell,
                double beta, double* K, int64_t ldk, int fill);

/* Symmetric InverseMultiquadric(Euclidean) kernel matrix: K[i,j] = (c2 + d2/ell^2)^(-1/2) (kernels.jl:156-164), 0 < c2, 0 < ell. */
int pl_gram_imq(const double* X, int64_t n, int64_t d, int64_t ldx, int cinv_kind, const double* cinv, double c2, double ell,
                double* K, int64_t ldk, int fill);

/* Rectangular kernel matrix KernelMatrix(F1,F2,k) (kernels.jl:229-244), written ROW-MAJOR n1 x n2:
 *   K[i*ldk + j] = k(X1_i, X2_j).   (Julia column-major m x n: pass F2 as X1 and F1 as X2.)
 * kernel_id: PL_KERNEL_LINEAR / DOT / RBF / IMQ; alpha used by DOT; cinv_kind, cinv, ell, beta by RBF (IMQ: beta = c2). */
int pl_gram_cross(int kernel_id, const double* X1, int64_t n1, int64_t ldx1, const double* X2, int64_t n2, int64_t ldx2,
                  int64_t d, int alpha, int cinv_kind, const double* cinv, double ell, double beta, double* K,
                  int64_t ldk);

/* Multi-GPU form of the symmetric kernel matrix for one process per GPU: this call computes only the 128x128 tiles
 * t with t % nparts == part of the upper triangle (no collective: tiles are independent, kernels.jl:217-221 loops
 * j >= i only) and returns them tile-packed, 128*128 doubles each, in pl_gram_tile_coords order.
 * tiles: pl_gram_tile_count(n, 0, part, nparts) * 16384 doubles. */
int pl_gram_part(int kernel_id, const double* X, int64_t n, int64_t d, int64_t ldx, int alpha, int cinv_kind, const double* cinv,
                 double ell, double beta, int part, int nparts, double* tiles);

/* ---- resident kernel matrix: KernelMatrix(F, k) (kernels.jl:211-223) too large for one host matrix -------------------------------------
 * N = 200 000 (BASELINE configs[4], hierarchical kDPP) has a 160 GB upper triangle: it only exists as tile-packed shards on the GPUs.
 * pl_gram_resident computes the 128 x 128 tiles of the upper triangle and LEAVES them on the devices: dealt round-robin over the devices
 * of pl_init_all, or (one process per GPU, collective) this rank's tiles t % nranks == rank. X is replicated; no collective on the result.
 * pl_gram_fetch_block gathers any block K[i0:i1, j0:j1] (row-major (i1-i0) x (j1-j0), both triangles by mirroring, <= 2 GB) to the host:
 * every shard writes the elements it owns into a zeroed staging block, one all-reduce assembles it (collective in the process-per-GPU
 * mode; every rank receives the block). pl_gram_resident_shard exposes a shard's device pointer (tile k of shard `part` is tile
 * pl_gram_tile_coords(n, 0, part, nparts, k)). The handle owns the device memory until pl_gram_free. */
int pl_gram_resident(int kernel_id, const double* X, int64_t n, int64_t d, int64_t ldx, int alpha, int cinv_kind, const double* cinv, double ell,
                     double beta, void** handle);
int pl_gram_resident_info(void* handle, int64_t* n, int* nparts, int64_t* tiles_local, int64_t* bytes_local);
int pl_gram_resident_shard(void* handle, int local_index, int* part, int64_t* ntiles, void** dtiles);
int pl_gram_fetch_block(void* handle, int64_t i0, int64_t i1, int64_t j0, int64_t j1, double* out, int64_t ldo);
int pl_gram_free(void* handle);

/* ---- weighted normal equations: replaces every A'WA / A'Wb site of src/Learning/linear-learn.jl ------- */
/* A: row-major R x Kf (rows = [n_energy_rows energy rows ; force rows], linear-learn.jl:232-245).
 * Row weights: w != NULL -> w[r] (ooc_learn! per-configuration weights, linear-learn.jl:341-352);
 *              w == NULL -> r < n_energy_rows ? w_e : w_f      (Q = Diagonal([ws1*1_N; ws2*1_M]), :246-249).
 * b: R right-hand sides or NULL (then AtWb is not touched).
 * add_intercept != 0 prepends the column [1_N ; 0_M] (linear-learn.jl:239-241): outputs have order Kf+1 and the
 * intercept is row/column 0. lambda is added to EVERY diagonal entry, intercept included (:253).
 * AtWA: (Kf+ic) x (Kf+ic) full symmetric (exactly symmetric), AtWb: Kf+ic. */
int pl_normal_eq(const double* A, int64_t R, int64_t Kf, int64_t lda, const double* w, int64_t n_energy_rows, double w_e,
                 double w_f, const double* b, int add_intercept, double lambda, double* AtWA, double* AtWb);

/* Streamed form over a LIST of host row blocks (direct ingestion of the reference's containers, no host-side
 * reduce(hcat/vcat) copy). blocks[k] is rows[k] x Kf row-major and dense (ld = Kf): lp.dB[i] (K x 3natoms column-major,
 * linear-learning-problem.jl:127) and reduce(hcat, lp.B) are passed as they lie in memory. The logical matrix is the concatenation of
 * the blocks in order; w and b index its R = sum(rows) rows; all other arguments as pl_normal_eq. Rows travel to the device in chunks
 * of chunk_rows (<= 0: library default, ~512 MB) through two alternating buffers while the SYRK of the previous chunk runs; chunk
 * partials are accumulated in chunk order (deterministic). Device footprint: two chunks, not R x Kf. */
int pl_normal_eq_blocks(const double* const* blocks, const int64_t* rows, int64_t nblocks, int64_t Kf, const double* w, int64_t n_energy_rows,
                        double w_e, double w_f, const double* b, int add_intercept, double lambda, int64_t chunk_rows, double* AtWA,
                        double* AtWb);

/* Out-of-core accumulation — ooc_learn! (linear-learn.jl:301-391) computes the descriptors of one configuration at a time and adds
 * A'WA, A'Wb to running sums (:355-356); the data never exists as one matrix. pl_neq_begin zeroes accumulators resident on the
 * device; every pl_neq_add contributes a batch of rows laid out [n_energy_rows energy rows ; force rows] with the weights of
 * pl_normal_eq (per-row w for the per-configuration energy weights ws1*{1, 1/natoms, 1/natoms^2}, :341-352), synchronously (the
 * host buffers may be reused on return); pl_neq_finish assembles (AtWA, AtWb) like pl_normal_eq and may be called any number of
 * times. Batches are summed in call order: deterministic for a fixed batching. */
int pl_neq_begin(int64_t Kf);
int pl_neq_add(const double* A, int64_t R, int64_t lda, const double* w, int64_t n_energy_rows, double w_e, double w_f, const double* b);
int pl_neq_finish(int add_intercept, double lambda, double* AtWA, double* AtWb);

/* One process per GPU (after pl_comm_init_rank; collective: every rank calls it): A_local is THIS rank's shard, laid out [its energy rows ;
 * its force rows]; the result is the normal equations of the row-wise union of all shards, identical on every rank. SYRK of the shard ->
 * ONE ncclAllReduce of the bucket [S ; v ; c ; s] (Kf^2 + 2 Kf + 2 doubles) -> finish, all on the library's stream. Other arguments as
 * pl_normal_eq. Without a communicator it equals pl_normal_eq. */
int pl_normal_eq_sharded(const double* A_local, int64_t R_local, int64_t Kf, int64_t lda, const double* w, int64_t n_energy_rows_local, double w_e,
                         double w_f, const double* b, int add_intercept, double lambda, double* AtWA, double* AtWb);
/* the same on DEVICE data, enqueue only (stream is a cudaStream_t): what the multi-GPU benchmark times */
int pl_normal_eq_sharded_dev(const double* dA, int64_t R_local, int64_t Kf, int64_t lda, const double* dw, int64_t n_energy_rows_local, double w_e,
                             double w_f, const double* db, int add_intercept, double lambda, double* dAtWA, double* dAtWb, void* stream);

/* ---- GlobalSum / GlobalMean features on the device ("next" row f2 of SURVEY.md 8f) ------------------------ */
/* locals: the local descriptors of all N configurations stacked, row-major (offsets[N]) x d, leading dimension ldl; configuration i owns
 * rows [offsets[i], offsets[i+1]) (offsets[0] == 0, strictly increasing). X[i,:] = sum of its rows (mean == 0: compute_feature(B,
 * ::GlobalSum) features.jl:19-21, sum.(get_values.(...)) linear-learning-problem.jl:75, pca_state.jl:23) or that sum / natoms_i
 * (mean != 0: ::GlobalMean, features.jl:32-34). Summation order is Julia's (left to right below 1024 atoms, pairwise above). */
int pl_global_feature(const double* locals, int64_t ldl, const int64_t* offsets, int64_t N, int64_t d, int mean, double* X, int64_t ldx);
int pl_global_feature_dev(const double* dL, int64_t ldl, const int64_t* doffsets, int64_t N, int64_t d, int mean, double* dX, int64_t ldx,
                          int64_t total_rows, void* stream);

/* CorrelationMatrix features (src/Kernels/features.jl:48-50, widened row f4): C[i] = mean(b*b' for b in the local descriptors of
 * configuration i), N matrices d x d (row-major, full symmetric), same stacked layout / offsets as pl_global_feature. Products and
 * sums are formed in the reference's order without fma contraction: bit-identical features. */
int pl_correlation_feature(const double* locals, int64_t ldl, const int64_t* offsets, int64_t N, int64_t d, double* C);

/* KernelSteinDiscrepancy with an RBF(Euclidean) kernel: compute_divergence(x, ::KernelSteinDiscrepancy) (src/Kernels/divergences.jl:27-49,
 * widened row f4). X: the n samples (row-major n x d), S: their scores div.score.(x) evaluated by the caller (n x d). The O(n^2 d)
 * pair loop becomes four Gram matrices on the DMMA engine plus one fused pair reduction; the reference's formulas are kept as written
 * (including the scalar broadcast inside compute_gradxy_kernel, kernels.jl:121-125). Needs 4 n^2 doubles of device workspace. */
int pl_ksd_rbf(const double* X, const double* S, int64_t n, int64_t d, int64_t ldx, int64_t lds, int cinv_kind, const double* cinv, double ell,
               double beta, double* ksd);

/* ---- covariance / second moment: replaces compute_eigen's Q (dimension_reduction.jl:11-12) and the
 *      mean/centre of fit!(ds, ::PCAState) (pca_state.jl:34-37) ------------------------------------------ */
/* X: row-major n x Kf.  center == 0: C = (1/n) sum_r x_r x_r'   (PCA / ActiveSubspace, pca.jl:28, as.jl:25).
 * center != 0: if mean_given == 0, mean_inout <- sum_r x_r / n (pca_state.jl:35), else mean_inout is used as is
 * (pca.m persists, pca_state.jl:34); C = (1/n) sum_r (x_r-m)(x_r-m)'. C: Kf x Kf full, exactly symmetric. */
int pl_cov(const double* X, int64_t n, int64_t Kf, int64_t ldx, double* mean_inout, int mean_given, int center, double* C);

/* One process per GPU (collective): X_local is this rank's n_local of the n_total rows. center != 0 and mean_given == 0: ONE pass over the rows
 * per rank (shifted second moment about the mean of a prefix of the shard), an all-reduce of the K column sums for the global mean
 * (returned in mean_inout on every rank), re-centring, ONE all-reduce of K x K. C = (1/n_total) sum over all ranks, exactly symmetric. */
int pl_cov_sharded(const double* X_local, int64_t n_local, int64_t n_total, int64_t Kf, int64_t ldx, double* mean_inout, int mean_given, int center,
                   double* C);
int pl_cov_sharded_dev(const double* dX, int64_t n_local, int64_t n_total, int64_t Kf, int64_t ldx, double* dmean_inout, int mean_given, int center,
                       double* dC, void* stream);
/* sum-all-reduce count doubles in place over the library's communicator, enqueued on stream (a no-op without a communicator) */
int pl_allreduce_dev(double* dbuf, int64_t count, void* stream);

/* ---- symmetric eigendecomposition and the L-ensemble on the device ("next" row f1 of SURVEY.md 8f) ----------- */
/* eigen(Symmetric(Q)) (dimension_reduction.jl:13; LAPACK in the reference) and the eigendecomposition inside EllEnsemble(K)
 * (Determinantal.jl [external], reached from dpp.jl:26,41).
 * n <= 2048: the library's OWN solver (csrc/eigh.cuh) -- Householder tridiagonalisation in one cooperative kernel (n <= 1184, the
 * sizes of configs 1, 3, 4 -- N = 1000, K = 500 / 1000: the matrix in registers, 2.7x faster than cuSOLVER; above: rows in L2 / shared memory, 1.6x),
 * divide and conquer on the tridiagonal matrix with its eigenvector updates on the DMMA Gram engine, blocked back-transformation.
 * No library call; results are bitwise reproducible.
 * Larger n: cuSOLVER's Xsyevd, loaded with dlopen at first use (PLB200_CUSOLVER overrides the search path).
 * PLB200_EIGH=own|cusolver forces one path (own: n <= 4096), PLB200_EIGH_OWN_MAX moves the switch-over.
 * Accuracy (own path, measured over random SPD / rank-deficient kernel / RBF / graded covariance / clustered / Wilkinson matrices up
 * to n = 4096): eigenvalues <= 1.5e-13 ||A|| (Wilkinson, n = 4096: LAPACK-level; <= 6e-14 up to n = 2048), residual |A U - U diag(W)| <= 7e-15 ||A||, |U'U - I| <= 3e-14.
 * Errors: PL_EINVAL for a matrix with Inf / NaN entries (Julia's eigen throws there too).
 * A: symmetric n x n, only the `fill` triangle of the row-major view is read. W: n eigenvalues, ascending. V (may be NULL: values only):
 * ROW k = unit eigenvector k, i.e. Julia's column-major `eigen(...).vectors` (column k) byte for byte. Signs are the solver's. */
int pl_eigh(const double* A, int64_t n, int64_t lda, int fill, double* W, double* V, int64_t ldv);
/* in place on device data: dA's rows become the eigenvectors; synchronises the stream (the solver reports through a device flag). */
int pl_eigh_dev(double* dA, int64_t n, int64_t lda, int fill, int want_vectors, double* dW, void* stream);

/* EllEnsemble(KernelMatrix(F, k)) (dpp.jl:25-26, 40-41) without the kernel matrix ever leaving the GPU: Gram kernel -> in-place
 * eigendecomposition. lambda: n eigenvalues (ascending). Kout (may be NULL): the kernel matrix in Symmetric(:U) storage (as pl_gram_*
 * with PL_FILL_JULIA_U; caller passes zeros(n,n)). U and lambda stay RESIDENT in the library until the next pl_ell_ensemble (or pl_finalize). */
int pl_ell_ensemble(int kernel_id, const double* X, int64_t n, int64_t d, int64_t ldx, int alpha, int cinv_kind, const double* cinv, double ell,
                    double beta, double* lambda, double* Kout, int64_t ldk);
/* inclusion_prob of the resident ensemble rescaled by alpha_scale (rescale!, dpp.jl:27): p[i] = sum_k U[i,k]^2 g_k,
 * g_k = a*lam_k/(1 + a*lam_k) (Determinantal.inclusion_prob, dpp.jl:69). */
int pl_ell_inclusion_prob(double alpha_scale, int64_t n, double* p);
/* rows idx[0..m) of the resident eigenvector matrix (the eigenvectors a k-DPP draw selected): V is m x n row-major. Only m*n doubles
 * cross PCIe instead of n*n. */
int pl_ell_eigvecs(const int64_t* idx, int64_t m, int64_t n, double* V, int64_t ldv);

/* First phase of a k-DPP draw (Determinantal.sample, dpp.jl:51 [external]; Kulesza & Taskar 2012 Alg. 7 + Alg. 8): which eigenvectors.
 * lam: the n eigenvalues already scaled by rescale!'s alpha (>= 0); uniforms: n numbers in [0, 1) from the CALLER's generator, uniforms[i]
 * decides eigenvalue n-1-i (visited from the largest index down, unused once k eigenvectors are taken). idx (capacity k): the selected
 * eigenvector indices (0-based, descending); *count: how many (k unless the table underflows). The (n+1) x (k+1) table of log elementary
 * symmetric polynomials is built on the device by one thread-block cluster (a cluster barrier per eigenvalue) and never leaves it. */
int pl_esp_select(const double* lam, int64_t n, int64_t k, const double* uniforms, int64_t* idx, int64_t* count);

/* Second phase of a k-DPP draw on the resident ensemble (Determinantal.sample, dpp.jl:51 [external]; Kulesza & Taskar 2012 Alg. 1):
 * projection-DPP sampling over the m eigenvectors eig_idx selected by the first phase. uniforms: m numbers in [0, 1) from the CALLER's
 * generator (the RNG stays in Julia / numpy), one per step; items: the m selected item indices (0-based, in draw order). U never leaves
 * the device. For n <= 32768 the conditioning runs in its kernel-row (pivoted Cholesky) form on P = V V', built once by the Gram engine:
 * a step reads t rows of an m x n factor; larger n use the O(n m^2) Gram-Schmidt form on V alone (PLB200_DPP_GS=1 forces it). Both pick
 * item t from the same conditional distribution by inverse CDF. */
int pl_ell_sample(const int64_t* eig_idx, int64_t m, int64_t n, const double* uniforms, int64_t* items);

/* Greedy MAP of the resident ensemble (Determinantal.greedy_subset, dpp.jl:60 [external]): k items, each maximising the conditional
 * variance given the ones before (fast greedy form: one row of K and a rank-t correction per step, O(n k^2) in total). The kernel matrix
 * is kept on the device next to its eigenvectors for n <= 32768. items[t] = -1 from the step at which the numerical rank is exhausted. */
int pl_ell_greedy(int64_t k, int64_t n, int64_t* items);

/* ---- DBSCAN's distance matrices between configurations (src/SubsetSelection/dbscan.jl:100-170, kabsch.jl; widened row f4) ---------- */
/* P: n configurations x natoms x 3 atomic positions (row-major; every configuration has natoms atoms). D: n x n, symmetric, zero diagonal.
 * pl_rmsd_periodic: distance_matrix_periodic / periodic_rmsd (minimum image in an orthorhombic box, box_lengths[3]; same operations in the
 * same order as the reference: bit-identical). pl_rmsd_kabsch: distance_matrix_kabsch / kabsch_rmsd (centre both, optimal rotation from the
 * 3 x 3 SVD of the covariance with the reflection check, RMSD of the superposition). One GPU thread per pair. */
int pl_rmsd_periodic(const double* P, int64_t n, int64_t natoms, const double* box_lengths, double* D, int64_t ldd);
int pl_rmsd_kabsch(const double* P, int64_t n, int64_t natoms, double* D, int64_t ldd);

/* ---- predictions of the fitted linear model ("next" row f3 of SURVEY.md 8f) ------------------------------ */
/* y[r] = A[r,:] . beta + (r < n_energy_rows ? beta0 : 0) over the same packed rows as pl_normal_eq: what
 * get_all_energies(ds, lb) / get_all_forces(ds, lb) (src/Data/utils.jl:42-65) evaluate after learn! (B.beta + beta0, dB.beta). */
int pl_predict(const double* A, int64_t R, int64_t Kf, int64_t lda, const double* beta, double beta0, int64_t n_energy_rows, double* y);
int pl_predict_dev(const double* dA, int64_t R, int64_t Kf, int64_t lda, const double* dbeta, double beta0, int64_t n_energy_rows,
                   double* dy, void* stream);

/* Residual metrics of predictions against reference values (src/Metrics/metrics.jl:11-53): out = [mae, rmse, rsq, mean_cos].
 * mae = sum|x_pred - x| / n, rmse = sqrt(sum (x_pred - x)^2 / n), rsq = 1 - sum (x_pred - x)^2 / sum (x - mean(x))^2 (two passes),
 * mean_cos = mean over the 3-vectors (forces, n % 3 == 0) of dot(x, x_pred) / (|x| |x_pred|) with NaNs filtered (NaN otherwise). */
int pl_metrics(const double* x_pred, const double* x, int64_t n, double* out);
int pl_metrics_dev(const double* dx_pred, const double* dx, int64_t n, double* dout, void* stream);

/* ---- device-resident entry points (enqueue only; `stream` is a cudaStream_t) ---------------------------- */

/* Kernel matrix on device data. dX2 == NULL -> symmetric (n2/ldx2 ignored, `fill` applies); otherwise rectangular
 * row-major n1 x n2 (fill ignored). part/nparts shard the 128x128 output tiles round-robin (tile t is computed iff
 * t % nparts == part); with packed != 0 the computed tiles are stored contiguously (tile-packed, 128*128 doubles
 * each, in the order given by pl_gram_tile_coords) instead of densely in an n1 x n2 matrix with leading dim ldo. */
int pl_gram_dev(int kernel_id, const double* dX1, int64_t n1, int64_t ldx1, const double* dX2, int64_t n2, int64_t ldx2,
                int64_t d, int alpha, int cinv_kind, const double* dcinv, double ell, double beta, double* dOut,
                int64_t ldo, int fill, int part, int nparts, int packed, void* stream);
/* number of tiles part `part` of `nparts` owns for an n1 x n2 (n2 = 0: symmetric n1 x n1) kernel matrix */
int64_t pl_gram_tile_count(int64_t n1, int64_t n2, int part, int nparts);
/* tile-row / tile-column (units of 128) of the k-th tile owned by `part`; returns 0 or PL_EINVAL */
int pl_gram_tile_coords(int64_t n1, int64_t n2, int part, int nparts, int64_t k, int64_t* ti, int64_t* tj);

/* The schedule pl_syrk_dev would use for an R x Kf matrix on a device with num_sms SMs, as a table (host arithmetic only; no device
 * needed): rows (cta, unit, first_stage, nstages, slot) — CTA `cta` streams `nstages` stages of 16 rows of work unit `unit` from stage
 * `first_stage` and writes its partial to workspace slot `slot`. mode 0: "waves" (one piece per CTA, ~20 short CTAs per SM);
 * mode 1: "streamk" (the cost-weighted work cut into num_sms x waves equal intervals, a CTA crossing unit boundaries without refilling
 * its pipeline). waves <= 0 / diag_cost <= 0: library defaults. table may be NULL (counts only). For tests and tools. */
int pl_syrk_schedule(int64_t R, int64_t Kf, int num_sms, int mode, int waves, double diag_cost, int64_t capacity, int* table, int64_t* npieces,
                     int64_t* nctas);

/* Partial weighted SYRK of a row shard: S = sum_r w_r (a_r-m)(a_r-m)' and, if dvec != NULL,
 * v = sum_r w_r b_r a_r, c = sum_{r<n_energy_rows} w_r a_r, s = [sum_{r<ne} w_r, sum_{r<ne} w_r b_r].
 * dmean == NULL -> no centring. Outputs are UNSCALED sums so shards can be all-reduced:
 *   dS  : Kf x Kf full symmetric, leading dim Kf
 *   dvec: 2*Kf + 2 doubles = [v (Kf) ; c (Kf) ; s (2)]  (may be NULL if neither b nor the intercept is wanted; not with dmean) */
int pl_syrk_dev(const double* dA, int64_t R, int64_t Kf, int64_t lda, const double* dw, int64_t n_energy_rows, double w_e,
                double w_f, const double* db, const double* dmean, double* dS, double* dvec, void* stream);
/* One-pass centred covariance of a row shard (no separate mean pass over the data): with a shift z close to the mean (e.g. the mean
 * of a prefix of the rows, pl_colsum_dev), pl_cov_shifted_dev returns S_z = sum_r (x_r - z)(x_r - z)' and rho = sum_r (x_r - z).
 * The shard's column sum is n_r z + rho (all-reduce it for the global mean m); pl_cov_recentre_dev then moves S_z to the mean:
 * S <- S - rho delta' - delta rho' + n_r delta delta', delta = m - z. Because z is within sigma/sqrt(prefix) of m nothing cancels. */
int pl_cov_shifted_dev(const double* dA, int64_t R, int64_t Kf, int64_t lda, const double* dz, double* dS, double* drho, void* stream);
int pl_cov_recentre_dev(double* dS, int64_t Kf, const double* drho, const double* ddelta, double n_rows, void* stream);
/* Assemble (AtWA, AtWb) of order Kf+ic from (all-reduced) S / vec as pl_normal_eq defines them. */
int pl_normal_eq_finish_dev(const double* dS, const double* dvec, int64_t Kf, int add_intercept, double lambda,
                            double* dAtWA, double* dAtWb, void* stream);
/* column sums of a row shard: dsum[k] = sum_r A[r,k] (deterministic two-stage reduction) */
int pl_colsum_dev(const double* dA, int64_t R, int64_t Kf, int64_t lda, double* dsum, void* stream);
/* dOut[i] = dIn[i] / denom, i < count (covariance 1/n and mean 1/n after the all-reduce) */
int pl_scale_div_dev(const double* dIn, double denom, int64_t count, double* dOut, void* stream);
/* dOut[i][j] = dOut[j][i] = dIn[min(i,j)][max(i,j)] / denom for a K x K matrix: the covariance 1/n after the all-reduce, made
 * EXACTLY symmetric again (a ring all-reduce sums S[i][j] and S[j][i] in different rank orders). */
int pl_sym_scale_div_dev(const double* dIn, int64_t K, double denom, double* dOut, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PLB200_H */
