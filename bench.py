#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 hot path (driver contract in the task statement).

Workload (BASELINE.json configs[1]): KernelMatrix(F, RBF(Euclidean(300))) for N = 20 000 GlobalMean descriptors of
length d = 300, fp64, synthetic (mu + N(0, 1/96)). One "step" = one full symmetric kernel matrix.

  value     : Gram GFLOP/s with the descriptors already resident in HBM; algorithmic flops N(N+1)d (symmetric-half
              convention, the reference itself only loops j >= i, kernels.jl:217-218) / device time of the step
              (CUDA events on the launching stream; the step = column-mean + row-norm preludes + the fused DMMA kernel).
  e2e       : the same metric through the host C ABI (pl_gram_rbf / pl_gram_part) with PINNED HOST buffers: H2D of the
              descriptors and D2H of the result inside the timed region.
  roofline  : the fused Gram kernel alone (events inside libplb200 around that launch) against the fp64 tensor (DMMA)
              peak, taken as cuBLAS DGEMM 8192^3 measured in this same process (MEASURED_PEAKS.json has no fp64 figure).
  N > 1     : one process per GPU; the 128x128 tiles of the upper triangle are dealt round-robin to ranks (no data-path
              collective), strong scaling of the one matrix; time = max over ranks.
  --impl reference : the reference's own algorithm (single-threaded scalar pair loop, kernels.jl:211-223) as restated in
              oracle/pl_oracle.c, timed on a bounded sample of rows of the same workload on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "potentiallearning.jl_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

N_CFG, D_CFG = 20000, 300
ELL, BETA = 1.0, 1.0
SEED = 1002
METRIC = "gram_rbf_fp64_gflops"
UNIT = "GFLOP/s"


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout. Libraries (NCCL's version banner, torchrun children, BLAS warnings) write to fd 1 too:
    keep a private duplicate of the real stdout for the line and point fd 1 at stderr for everything else."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def make_features_host(n=N_CFG, d=D_CFG, seed=SEED):
    rng = np.random.default_rng(seed)
    mu = 2.0 * rng.standard_normal(d)
    return mu + rng.standard_normal((n, d)) / np.sqrt(96.0)


def workload_config(extra=None):
    cfg = {"workload": f"KernelMatrix RBF(Euclidean({D_CFG})) N={N_CFG} d={D_CFG} (BASELINE configs[1])", "flop_convention": "N(N+1)d",
           "l2": "output 3.2 GB >> 126 MB L2; 256 MiB L2 flush written between timed GPU steps"}
    if extra:
        cfg.update(extra)
    return cfg


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region. NVML (pynvml) is polled every ~5 ms in a thread; if NVML is not
    importable the recipe's nvidia-smi query is used instead (slower: ~50 ms per sample)."""
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index = index
        self.sm, self.reason_bits, self.max_sm, self.power = [], 0, None, []
        self.smi_samples = []
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    @staticmethod
    def _physical_index(i):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[i])
            except Exception:
                return i
        return i

    def _run(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                if nv is not None:
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
                    try:
                        self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                    except Exception:
                        self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                    self._stop.wait(0.005)
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.smi_samples.append([x.strip() for x in out.split(",")])
                    self._stop.wait(0.05)
            except Exception:
                self._stop.wait(0.05)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if self._nvml is not None and self.sm:
            nv = self._nvml
            names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            reasons = [k for k, bit in names.items() if self.reason_bits & bit]
            return {"sm_mhz": float(np.median(self.sm)), "sm_min_mhz": float(min(self.sm)), "sm_max_mhz": self.max_sm, "reasons": reasons,
                    "power_w_max": max(self.power) if self.power else None, "samples": len(self.sm), "source": "nvml"}
        if not self.smi_samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(s[0]) for s in self.smi_samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.smi_samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 3 + i and s[3 + i].lower().startswith("active") for s in self.smi_samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.smi_samples), "source": "nvidia-smi"}


def cpu_reference_rate(X, seconds_target, oracle):
    """Time the restated reference pair loop on the first rows of the upper triangle; returns (GFLOP/s, pairs, seconds, rows)."""
    k = oracle.RBF(oracle.Euclidean(X.shape[1]), ell=ELL, beta=BETA)
    n, d = X.shape
    t0 = time.perf_counter()
    pairs = oracle.kernel_matrix_rows(X, k, 0, 8)  # calibrate
    dt = time.perf_counter() - t0
    rows = int(max(8, min(n, 8 * seconds_target / max(dt, 1e-6))))
    t0 = time.perf_counter()
    pairs = oracle.kernel_matrix_rows(X, k, 0, rows)
    dt = time.perf_counter() - t0
    return pairs * 2.0 * d / dt / 1e9, pairs, dt, rows


def cpu_strong_blas_rate(X, n_sub=6000):
    """What a well-written CPU version would do (NOT what the reference does): OpenBLAS dsyrk on all host threads plus a
    vectorised distance/exp epilogue, on the first n_sub rows. Reported next to the faithful port for context only."""
    from scipy.linalg import blas

    Xs = np.ascontiguousarray(X[:n_sub])
    n, d = Xs.shape
    t0 = time.perf_counter()
    Xc = Xs - Xs.mean(0)
    G = blas.dsyrk(1.0, Xc, lower=0, trans=0)  # upper triangle of Xc Xc'
    q = np.einsum("ij,ij->i", Xc, Xc)
    D2 = q[:, None] + q[None, :] - 2.0 * G
    np.maximum(D2, 0.0, out=D2)
    K = BETA * np.exp(-0.5 * D2 / ELL ** 2)
    dt = time.perf_counter() - t0
    del K
    return n * (n + 1) * d / dt / 1e9, dt


def other_configs(dev, peak_tflops, rank, world, sync):
    """The other BASELINE shapes at this N, inputs resident in HBM, CUDA events on the launching stream, best of `reps` after warm-ups,
    max over ranks per repetition. Every rank calls this (the sharded entries are collective); rank 0 reports.
      configs[2] normal equations: rows sharded (whole configurations per rank), fused SYRK -> ONE ncclAllReduce of the bucket
                 [S ; v ; c ; s] issued by libplb200 itself (pl_normal_eq_sharded_dev) -> finish kernel; rows/s over ALL ranks;
      configs[3] PCA covariance of 10^7 x 1000 rows (80 GB over the ranks; fewer rows if one GPU cannot hold its share): one pass,
                 two all-reduces (K column sums, K x K) inside pl_cov_sharded_dev;
      configs[4] DotProduct Gram N = 200 000 x 500, tile-packed shards (tiles round-robin over ranks, no collective);
      configs[0] DotProduct Gram N = 1000 x 26 (the kDPP example shape), rank 0 only: HBM-bound, reported against the copy peak.
    Reported next to the headline metric, not instead of it."""
    import torch
    import torch.distributed as dist

    from plb200 import _lib

    lib = _lib.lib()
    res = {"comm": _lib.comm_info() if world > 1 else None}
    if world > 1:
        res["comm"]["exchange"] = ("peer memory: reduce-scatter + all-gather over NVLink loads/stores in one libplb200 kernel behind the SYRK's reduce kernel"
                                   if res["comm"]["peer_exchange"] else "ncclAllReduce issued by libplb200")
    stream = lambda: torch.cuda.current_stream(dev).cuda_stream

    def timed(fn, warm=2, reps=4):
        for _ in range(warm):
            fn()
        sync()
        ts, main, coll = [], [], []
        _lib.set_profiling(True)
        for _ in range(reps):
            sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            t = torch.tensor([e0.elapsed_time(e1), _lib.last_main_kernel_ms(), _lib.last_coll_ms() if world > 1 else 0.0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t[0]))
            main.append(float(t[1]))
            coll.append(float(t[2]))
        _lib.set_profiling(False)
        i = int(np.argmin(ts))
        return ts[i], main[i], coll[i]

    def free_gb():
        torch.cuda.empty_cache()
        return torch.cuda.mem_get_info(dev)[0] / 1e9

    try:
        # ---- configs[2]: learn! normal equations, 10 000 configurations x 96 atoms, K = 500 ----
        g = torch.Generator(device=dev).manual_seed(1003 + rank)
        ncfg, nat, K = 10_000, 96, 500
        lo, hi = rank * ncfg // world, (rank + 1) * ncfg // world
        ne_l = hi - lo
        R_l = ne_l * (1 + 3 * nat)
        R = ncfg * (1 + 3 * nat)
        A = torch.empty((R_l, K), device=dev, dtype=torch.float64)
        for r0 in range(0, R_l, 1 << 19):
            A[r0:r0 + (1 << 19)].normal_(generator=g)
        b = torch.randn(R_l, device=dev, dtype=torch.float64, generator=g)
        M = torch.empty((K + 1, K + 1), device=dev, dtype=torch.float64)
        v = torch.empty(K + 1, device=dev, dtype=torch.float64)

        def neq():
            _lib.check(lib.pl_normal_eq_sharded_dev(A.data_ptr(), R_l, K, K, None, ne_l, 100.0, 1.0, b.data_ptr(), 1, 0.01, M.data_ptr(), v.data_ptr(),
                                                    stream()), "pl_normal_eq_sharded_dev")

        ms, main_ms, coll_ms = timed(neq)
        tf = R * K * (K + 1.0) / (ms * 1e-3) / 1e12
        res["normal_equations_c3"] = {
            "workload": f"learn! energy+force normal equations R={R} K={K}, ws=[100,1], intercept, lambda=0.01 (BASELINE configs[2]); rows sharded over {world} GPU(s), ONE exchange of K^2+2K+2 doubles inside libplb200 (see comm.exchange)",
            "ms": ms, "rows_per_s": R / (ms * 1e-3), "tflops_RK(K+1)": tf, "frac_of_measured_dgemm_per_gpu": tf / world / peak_tflops,
            "syrk_kernel_ms_max_rank": main_ms, "exchange_ms_max_rank": coll_ms, "rows_per_rank": R_l}
        del A, b, M, v

        # ---- configs[3]: PCAState covariance, 10^7 rows x 1000 features ----
        n_total, K = 10_000_000, 1000
        n_l = n_total // world
        room = free_gb() - 8.0
        full = n_l * K * 8 / 1e9 <= room
        if not full:
            n_l = int(room * 1e9 / (K * 8)) // 65536 * 65536
            n_total = n_l * world
        g = torch.Generator(device=dev).manual_seed(1004)  # the same feature means on every rank
        shift = 10.0 * torch.randn(K, device=dev, dtype=torch.float64, generator=g)
        g = torch.Generator(device=dev).manual_seed(2004 + rank)
        X = torch.empty((n_l, K), device=dev, dtype=torch.float64)
        for r0 in range(0, n_l, 1 << 18):
            X[r0:r0 + (1 << 18)].normal_(generator=g)
            X[r0:r0 + (1 << 18)] += shift
        C = torch.empty((K, K), device=dev, dtype=torch.float64)
        m = torch.empty(K, device=dev, dtype=torch.float64)

        def cov():
            _lib.check(lib.pl_cov_sharded_dev(X.data_ptr(), n_l, n_total, K, K, m.data_ptr(), 0, 1, C.data_ptr(), stream()), "pl_cov_sharded_dev")

        ms, main_ms, coll_ms = timed(cov, warm=1, reps=3)
        tf = n_total * K * (K + 1.0) / (ms * 1e-3) / 1e12
        res["pca_covariance_c4"] = {
            "workload": f"PCAState fit! centred covariance, n={n_total} rows x K={K} (BASELINE configs[3]" + ("" if full else f"; REDUCED from 10^7 rows: {room:.0f} GB free per GPU") + f"); rows sharded over {world} GPU(s), one pass, exchanges of K and K^2 doubles inside libplb200",
            "ms": ms, "rows_per_s": n_total / (ms * 1e-3), "tflops_nK(K+1)": tf, "frac_of_measured_dgemm_per_gpu": tf / world / peak_tflops,
            "syrk_kernel_ms_max_rank": main_ms, "last_exchange_ms_max_rank": coll_ms, "rows_per_rank": n_l, "full_size": bool(full)}
        del X, C, m, shift

        # ---- configs[4]: DotProduct Gram for hierarchical kDPP, N = 200 000 x 500, tile-packed shards ----
        n5, d5 = 200_000, 500
        rng = np.random.default_rng(1005)
        mu = torch.as_tensor(2.0 * rng.standard_normal(d5), device=dev)
        g = torch.Generator(device=dev).manual_seed(1005)  # X is replicated: the same rows on every rank
        X5 = mu + torch.randn((n5, d5), device=dev, dtype=torch.float64, generator=g) / np.sqrt(96.0)
        passes = 1
        while int(lib.pl_gram_tile_count(n5, 0, rank, world * passes)) * 131072 / 1e9 > free_gb() - 6.0 and passes < 64:
            passes *= 2
        nt = int(lib.pl_gram_tile_count(n5, 0, 0, world * passes))  # part 0 owns the most tiles
        tiles = torch.empty((nt, 128, 128), device=dev, dtype=torch.float64)

        def gram5():
            for ps in range(passes):  # passes > 1 only when one GPU cannot hold its whole shard: later passes overwrite the buffer
                _lib.check(lib.pl_gram_dev(_lib.PL_KERNEL_DOT, X5.data_ptr(), n5, d5, None, 0, 0, d5, 2, 0, None, 1.0, 1.0, tiles.data_ptr(), 128,
                                           _lib.PL_FILL_FULL, rank * passes + ps, world * passes, 1, stream()), "pl_gram_dev")

        ms, main_ms, _ = timed(gram5, warm=1, reps=3)
        fl = float(n5) * (n5 + 1) * d5
        tf = fl / (ms * 1e-3) / 1e12
        res["gram_dot_c5"] = {
            "workload": f"KernelMatrix DotProduct N={n5} d={d5} (BASELINE configs[4]); 128x128 tiles of the upper triangle round-robin over {world} GPU(s), tile-packed shards resident ({nt * 131072 / 1e9:.1f} GB per GPU" + (f", computed in {passes} passes over one buffer" if passes > 1 else "") + "), no collective",
            "ms": ms, "gflops_N(N+1)d": fl / (ms * 1e-3) / 1e9, "tflops": tf, "frac_of_measured_dgemm_per_gpu": tf / world / peak_tflops, "passes": passes}
        del X5, tiles

        # ---- configs[0]: the kDPP example shape, N = 1000 x 26 DotProduct (HBM / launch bound), rank 0 ----
        if rank == 0:
            n1, d1 = 1000, 26
            X1 = torch.as_tensor(rng.standard_normal((n1, d1)) + 2.0 * rng.standard_normal(d1), device=dev)
            K1 = torch.zeros((n1, n1), device=dev, dtype=torch.float64)

            def gram1():
                _lib.check(lib.pl_gram_dev(_lib.PL_KERNEL_DOT, X1.data_ptr(), n1, d1, None, 0, 0, d1, 2, 0, None, 1.0, 1.0, K1.data_ptr(), n1,
                                           _lib.PL_FILL_JULIA_U, 0, 1, 0, stream()), "pl_gram_dev")

            for _ in range(5):
                gram1()
            torch.cuda.synchronize()
            reps = 200
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                gram1()
            e1.record()
            e1.synchronize()
            us = e0.elapsed_time(e1) / reps * 1e3
            wbytes = n1 * (n1 + 1) / 2 * 8 + n1 * d1 * 8
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
            hbm = float(peaks.get("hbm_gbs", 6545.0))
            res["gram_dot_c1"] = {"workload": f"kDPP kernel build: KernelMatrix DotProduct N={n1} d={d1} (BASELINE configs[0]), Symmetric(:U) storage, back-to-back launches (L2-resident: 8 MB result)",
                                  "us_per_matrix": us, "algorithmic_GBps": wbytes / (us * 1e-6) / 1e9, "hbm_peak_GBps": hbm,
                                  "frac_of_hbm_copy_peak": wbytes / (us * 1e-6) / 1e9 / hbm}
            # ---- the eigendecomposition behind EllEnsemble(K) (config 1, N = 1000) and compute_eigen (config 4, K = 1000): the
            # built-in solver (csrc/eigh.cuh) next to cuSOLVER Xsyevd on the same kernel matrix, device-resident, vectors wanted
            def eig_ms(mode):
                os.environ["PLB200_EIGH"] = mode
                ts = []
                for rep in range(5):
                    A = K1 + K1.T - torch.diag(torch.diagonal(K1))
                    W = torch.empty(n1, device=dev, dtype=torch.float64)
                    torch.cuda.synchronize()
                    e0.record()
                    _lib.check(lib.pl_eigh_dev(A.data_ptr(), n1, n1, _lib.PL_FILL_FULL, 1, W.data_ptr(), stream()), "pl_eigh_dev")
                    e1.record()
                    e1.synchronize()
                    ts.append(e0.elapsed_time(e1))
                os.environ.pop("PLB200_EIGH", None)
                return float(np.median(ts[1:])), W, A

            own_ms, W_own, U_own = eig_ms("own")
            lib_ms, W_lib, _ = eig_ms("cusolver")
            Kfull = (K1 + K1.T - torch.diag(torch.diagonal(K1)))
            resid = float((Kfull @ U_own.T - U_own.T * W_own[None, :]).abs().max() / W_own.abs().max())
            orth = float((U_own @ U_own.T - torch.eye(n1, device=dev, dtype=torch.float64)).abs().max())
            Kh = Kfull.cpu().numpy()
            t0 = time.perf_counter()
            np.linalg.eigh(Kh)
            lapack_ms = (time.perf_counter() - t0) * 1e3  # what the reference does here: LAPACK on the host cores
            res["eigh_c1"] = {"workload": f"eigen(Symmetric(K)) of the config-1 kernel matrix (N={n1}; the same size as compute_eigen at K=1000, configs[3]): built-in "
                                          "tridiagonalisation + divide and conquer + back-transformation, device-resident, eigenvectors wanted",
                              "ms": own_ms, "cusolver_xsyevd_ms_same_box": lib_ms, "speedup_vs_cusolver": lib_ms / own_ms,
                              "host_lapack_eigh_ms": lapack_ms, "host_cores": os.cpu_count(),
                              "max_eig_diff_vs_cusolver_rel": float((W_own - W_lib).abs().max() / W_lib.abs().max()),
                              "residual_rel": resid, "orthogonality": orth}
            # ---- config 1 end to end through the public host API: kDPP(features, DotProduct) = KernelMatrix + EllEnsemble(K) + rescale! +
            # inclusion probabilities (dpp.jl:35-44, 69), host features in, lambda / K / probabilities out; K (8 MB) travels back once
            import plb200 as _pl

            Xh = X1.cpu().numpy()
            ts = []
            for rep in range(4):
                t0 = time.perf_counter()
                ell = _pl.EllEnsemble.from_features(Xh, _pl.DotProduct())
                _pl.rescale_(ell, 200)  # below the numerical rank of the ensemble (d(d+1)/2 = 351 for DotProduct(2) on 26 features)
                pr = _pl.inclusion_prob(ell)
                ts.append((time.perf_counter() - t0) * 1e3)
            res["kdpp_c1_e2e"] = {"workload": f"kDPP(features, DotProduct(2)) N={n1} d={d1} batch_size=200 (BASELINE configs[0]) through the host API: kernel matrix, "
                                              "L-ensemble eigendecomposition (built-in solver), rescale!, inclusion probabilities; host arrays in and out",
                                  "ms": float(np.median(ts[1:])), "sum_inclusion_prob": float(pr.sum()),
                                  "host_lapack_eigh_alone_ms": lapack_ms}
            del X1, K1
    except Exception as e:  # never let the side measurements take the headline line down
        import traceback

        res["error"] = repr(e) + " | " + traceback.format_exc(limit=2).replace("\n", " / ")
    torch.cuda.empty_cache()
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle

    oracle.build_c()
    X = make_features_host()
    rates = []
    for _ in range(args.warmup):
        cpu_reference_rate(X, 1.0, oracle)
    t_total = 0.0
    sample = None
    for _ in range(args.steps):
        r, pairs, dt, rows = cpu_reference_rate(X, 3.0, oracle)
        rates.append(r)
        t_total += dt
        sample = f"rows 0..{rows} of the upper-triangle pair loop ({pairs} kernel evaluations of {N_CFG * (N_CFG + 1) // 2}) per step"
    value = float(np.mean(rates))
    full_ms = N_CFG * (N_CFG + 1) * D_CFG / (value * 1e9) * 1e3
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": full_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(), "arm_notes": {"ms_per_step": "extrapolated from the sampled rows to the full matrix"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": sample + "; the reference loop is single-threaded Julia (kernels.jl:217-221), restated in C (oracle/pl_oracle.c)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the extra (non-headline) BASELINE shapes reported under 'also'")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import plb200
    from plb200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # (NCCL's log goes to fd 1, which claim_stdout() has already pointed at stderr: NCCL_DEBUG stays whatever the launcher set)
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    plb200.init(local_rank)
    if world > 1:
        _lib.comm_init_from_torch()  # the library's own NCCL communicator: the data-path all-reduces are issued inside libplb200
    W = max(args.warmup, 3)
    K = args.steps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- inputs resident in HBM ----------------
    Xh = make_features_host()
    X = torch.as_tensor(Xh, device=dev)
    kern = plb200.RBF(plb200.Euclidean(D_CFG), ell=ELL, beta=BETA)
    packed = world > 1
    if packed:
        ntiles_mine = int(_lib.load().pl_gram_tile_count(N_CFG, 0, rank, world))
        out = torch.empty((ntiles_mine, 128, 128), device=dev, dtype=torch.float64)
    else:
        out = torch.zeros((N_CFG, N_CFG), device=dev, dtype=torch.float64)  # zeros(n,n) (kernels.jl:216)
    flush = torch.empty(256 * 1024 * 1024 // 8, device=dev, dtype=torch.float64)

    def step():
        # dense output: the reference's storage contract, Symmetric(K) with uplo = :U over zeros(n,n) (kernels.jl:216-222):
        # one triangle is computed and stored, exactly what the reference computes and stores
        plb200.kernel_matrix_dev(X, kern, out=out, fill=_lib.PL_FILL_JULIA_U, part=rank, nparts=world, packed=packed)

    flops = float(N_CFG) * (N_CFG + 1) * D_CFG
    for _ in range(W):
        step()
    barrier()
    launches0 = plb200.launch_count()
    _lib.set_profiling(True)
    step_ms, main_ms = [], []
    with ClockSampler(local_rank) as clk:
        barrier()
        t_wall0 = time.perf_counter()
        for _ in range(K):
            flush.fill_(1.0)  # L2 flush, outside the per-step events
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            e1.synchronize()
            step_ms.append(e0.elapsed_time(e1))
            main_ms.append(_lib.last_main_kernel_ms())
        barrier()
        t_wall = time.perf_counter() - t_wall0
    _lib.set_profiling(False)
    launches = plb200.launch_count() - launches0
    total_ms = torch.tensor([sum(step_ms)], device=dev, dtype=torch.float64)
    main_total = torch.tensor([sum(main_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(main_total, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / K
    main_ms_per_step = float(main_total.item()) / K
    value = flops / (ms_per_step * 1e-3) / 1e9

    # ---------------- end to end through the host C ABI, pinned host buffers ----------------
    e2e = None
    if not args.no_e2e:
        Xp = torch.as_tensor(Xh).pin_memory()
        lib = _lib.lib()
        if packed:
            Kp = torch.empty((ntiles_mine, 128, 128), dtype=torch.float64).pin_memory()
            d2h = Kp.numel() * 8

            def e2e_step():
                _lib.check(lib.pl_gram_part(_lib.PL_KERNEL_RBF, _lib.dptr(Xp.numpy()), N_CFG, D_CFG, D_CFG, 0, 0, None, ELL, BETA, rank, world,
                                            _lib.dptr(Kp.numpy().reshape(-1))), "pl_gram_part")
        else:
            Kp = torch.zeros((N_CFG, N_CFG), dtype=torch.float64).pin_memory()  # zeros(n,n) (kernels.jl:216)
            d2h = sum(min(r0 + 128, N_CFG) * (min(r0 + 128, N_CFG) - r0) * 8 for r0 in range(0, N_CFG, 128))

            def e2e_step():
                # Symmetric(K) storage contract of the reference: uplo = :U triangle only (kernels.jl:217-222)
                _lib.check(lib.pl_gram_rbf(_lib.dptr(Xp.numpy()), N_CFG, D_CFG, D_CFG, 0, None, ELL, BETA, _lib.dptr(Kp.numpy().reshape(-1)), N_CFG,
                                           _lib.PL_FILL_JULIA_U), "pl_gram_rbf")
        for _ in range(3):
            e2e_step()
        barrier()
        n_e2e = max(3, min(K, 10))
        t0 = time.perf_counter()
        per_call = []
        for _ in range(n_e2e):
            tc = time.perf_counter()
            e2e_step()  # synchronous: returns after the D2H has completed
            per_call.append((time.perf_counter() - tc) * 1e3)
        t_loop = time.perf_counter() - t0
        barrier()
        t_e2e = torch.tensor([(time.perf_counter() - t0) / n_e2e], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        tm = _lib.last_timing()
        e2e = {"value": flops / float(t_e2e.item()) / 1e9, "unit": UNIT, "h2d_bytes_per_step": N_CFG * D_CFG * 8, "d2h_bytes_per_step": int(d2h),
               "ms_per_step": float(t_e2e.item()) * 1e3, "steps": n_e2e, "last_call_ms": tm, "host_buffers": "pinned",
               "rank0_wall_ms_per_call": [round(x, 3) for x in per_call], "rank0_loop_ms_per_call": t_loop / n_e2e * 1e3,
               "api": "pl_gram_part (tile-packed shard)" if packed else "pl_gram_rbf fill=PL_FILL_JULIA_U (Symmetric uplo=:U storage)"}
        if world > 1:
            # where the slowest rank's time goes: kernel / H2D / D2H of every rank's last call (GPUs behind one PCIe switch share its uplink)
            mine = torch.tensor([tm["kernel_ms"], tm["h2d_ms"], tm["d2h_ms"], tm["host_wall_ms"]], device=dev, dtype=torch.float64)
            allr = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            e2e["last_call_ms_per_rank"] = [[round(float(x), 3) for x in t] for t in allr]
        del Kp
        if world == 1:
            # What the Julia shim does by default: the result is an ordinary (pageable) zeros(n, n). The library stages it through its own
            # pinned ring + host threads. "touched": the same array again (pages resident); "fresh": a new np.zeros each call (first touch).
            Xg = np.array(Xh)
            Kg = np.zeros((N_CFG, N_CFG))

            def pageable_step(Kdst):
                _lib.check(lib.pl_gram_rbf(_lib.dptr(Xg), N_CFG, D_CFG, D_CFG, 0, None, ELL, BETA, _lib.dptr(Kdst.reshape(-1)), N_CFG,
                                           _lib.PL_FILL_JULIA_U), "pl_gram_rbf")

            pageable_step(Kg)
            t0 = time.perf_counter()
            for _ in range(3):
                pageable_step(Kg)
            t_touched = (time.perf_counter() - t0) / 3
            t_fresh = []
            for _ in range(2):
                Kf = np.zeros((N_CFG, N_CFG))
                t0 = time.perf_counter()
                pageable_step(Kf)
                t_fresh.append(time.perf_counter() - t0)
                del Kf
            e2e["pageable"] = {"touched_ms_per_step": t_touched * 1e3, "touched_value": flops / t_touched / 1e9, "fresh_zeros_ms_per_step": min(t_fresh) * 1e3,
                               "fresh_zeros_value": flops / min(t_fresh) / 1e9,
                               "note": "pageable numpy result (what Julia's zeros(n,n) is): staged by libplb200's pinned ring + host threads"}
            del Kg, Xg

    # ---------------- roofline denominator: fp64 tensor peak = cuBLAS DGEMM measured here (every rank: `also` reports against it) ----------
    del out
    torch.cuda.empty_cache()
    a = torch.randn((8192, 8192), device=dev, dtype=torch.float64)
    b = torch.randn((8192, 8192), device=dev, dtype=torch.float64)
    c = torch.empty((8192, 8192), device=dev, dtype=torch.float64)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    peak_tflops = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
    del a, b, c

    # ---------------- the other BASELINE shapes at this N, device-resident, collectives inside the library (every rank takes part) ----------
    also = None
    if not args.no_also:
        _lib.check(_lib.lib().pl_workspace_trim(0), "pl_workspace_trim")
        del flush
        torch.cuda.empty_cache()
        also = other_configs(dev, peak_tflops, rank, world, barrier)

    if rank != 0:
        if world > 1:
            _lib.comm_destroy()
            dist.destroy_process_group()
        return 0

    flops_launch = flops / world  # tiles are dealt evenly; per-launch algorithmic flops of this rank
    achieved = flops_launch / (main_ms_per_step * 1e-3) / 1e12
    traffic, traffic_source = None, None
    for name in ("r02c_gram_rbf_ncu_summary.json", "r02_gram_rbf_ncu_summary.json", "r01_gram_rbf_ncu_summary.json"):
        prof = os.path.join(ROOT, "profiles", name)
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
                traffic_source = f"static: one ncu --set full capture of this kernel at this shape, committed as profiles/{name} (not measured in this run)"
                break
            except Exception:
                traffic = None
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops, "traffic": traffic,
                "traffic_source": traffic_source, "kernel": "plb::gram_kernel<EPI_RBF>", "kernel_ms": main_ms_per_step,
                "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul fp64) measured in this run; MEASURED_PEAKS.json has no fp64 figure; nominal 37 TFLOP/s (HGX B200)",
                "algorithmic_flops_per_launch": flops_launch}

    # ---------------- CPU baseline on the box's host cores (rank 0, N = 1 only, bounded sample) ----------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle

        oracle.build_c()
        r, pairs, dt, rows = cpu_reference_rate(Xh, 12.0, oracle)
        cpu = {"value": r, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"rows 0..{rows} of the reference pair loop ({pairs} of {N_CFG * (N_CFG + 1) // 2} kernel evaluations, {dt:.1f} s); "
                         "single thread because the reference loop is serial Julia (kernels.jl:217-221)", "host_cores_available": os.cpu_count()}
        try:
            sr, sdt = cpu_strong_blas_rate(Xh)
            cpu["strong_blas_context"] = {"value": sr, "unit": UNIT, "cores": os.cpu_count(),
                                          "what": f"OpenBLAS dsyrk + vectorised exp on the first 6000 rows ({sdt:.2f} s) — a BLAS-3 reformulation the reference does not have"}
        except Exception as e:  # scipy missing on the box: the faithful port above is the baseline
            cpu["strong_blas_context"] = {"unavailable": str(e)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            # `config` is the workload and nothing else: identical, key for key, in both arms (the driver compares them)
            "config": workload_config(),
            "arm_notes": {"parallelism": f"tiles round-robin over {world} GPU(s), no collective",
                          "output": "tile-packed shards" if packed else "dense N x N storage, Symmetric uplo=:U triangle written (reference contract, kernels.jl:216-222)"},
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clk.summary(),
            "wall_s_timed_region": t_wall, "also": also}
    emit(line)
    if world > 1:
        _lib.comm_destroy()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
