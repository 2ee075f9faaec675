"""Generates tests/golden/*.npz: small seeded inputs of the hot path together with an INDEPENDENT extended-precision
(np.longdouble, 80-bit x87) evaluation of the reference formulas. Run from the repo root:

    python tests/golden/make_golden.py

The reference itself (Julia) cannot be executed in this image or on the GPU box, and its own tests hold no golden values
for this path, so these vectors pin the ORACLE (oracle/oracle.py, oracle/pl_oracle.c) against a second,
differently-written evaluation of the same reference lines — not against a run of the reference ("parity unpinned").
Shapes follow the reference's tests: d = 8, 20 atoms, 10 / 8 / 100 configurations (test/kernels/kernel_tests.jl:7-9,
test/learning/linear_tests.jl:20-22,71-73), plus a slice of BASELINE config 1 (26-dim ACE-like descriptors).
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LD = np.longdouble


def truth_kernels(X, alpha, cinv, ell, beta):
    """DotProduct (kernels.jl:35) and RBF-Euclidean (kernels.jl:73-74, distances.jl:59) in extended precision."""
    Xl = X.astype(LD)
    n = X.shape[0]
    Kd = np.zeros((n, n), dtype=LD)
    Kr = np.zeros((n, n), dtype=LD)
    C = None if cinv is None else (np.diag(cinv.astype(LD)) if cinv.ndim == 1 else cinv.astype(LD))
    for i in range(n):
        for j in range(n):
            a, b = Xl[i], Xl[j]
            Kd[i, j] = (np.sum(a * b) / np.sqrt(np.sum(a * a) * np.sum(b * b))) ** alpha
            v = a - b
            d2 = np.sum(v * v) if C is None else v @ C @ v
            Kr[i, j] = LD(beta) * np.exp(LD(-0.5) * d2 / LD(ell) ** 2)
    return Kd.astype(np.float64), Kr.astype(np.float64)


def truth_normal_eq(A, b, q, lam):
    Al, bl, ql = A.astype(LD), b.astype(LD), q.astype(LD)
    M = Al.T @ (ql[:, None] * Al) + LD(lam) * np.eye(A.shape[1], dtype=LD)
    v = Al.T @ (ql * bl)
    return M.astype(np.float64), v.astype(np.float64)


def truth_cov(X, center):
    Xl = X.astype(LD)
    n = X.shape[0]
    m = Xl.sum(axis=0) / n
    Y = Xl - m if center else Xl
    return m.astype(np.float64), ((Y.T @ Y) / n).astype(np.float64)


def main():
    rng = np.random.default_rng(20261019)
    out = {}
    # kernel_tests.jl fixture: 10 configs x 20 atoms x d = 8 -> GlobalMean features
    ld = rng.standard_normal((10, 20, 8))
    feats = np.stack([np.add.reduce(c, axis=0) / 20 for c in ld])
    out["kt_local"] = ld
    out["kt_feat"] = feats
    out["kt_dot"], out["kt_rbf"] = truth_kernels(feats, 2, None, 1.0, 1.0)
    # config-1 slice: 48 configs x 32 atoms x 26 dims, mean-heavy
    mu = 2.0 * rng.standard_normal(26)
    f1 = mu + rng.standard_normal((48, 26)) / np.sqrt(32)
    cd = rng.uniform(0.5, 2.0, size=26)
    B = rng.standard_normal((26, 26))
    cf = B @ B.T / 26 + np.eye(26)
    out["c1_feat"], out["c1_cdiag"], out["c1_cfull"] = f1, cd, cf
    out["c1_dot3"], out["c1_rbf"] = truth_kernels(f1, 3, None, 0.8, 2.0)
    _, out["c1_rbf_diag"] = truth_kernels(f1, 2, cd, 1.5, 1.0)
    _, out["c1_rbf_full"] = truth_kernels(f1, 2, cf, 2.0, 1.0)
    # linear_tests.jl covariate fixture at reduced size: 12 configs x 5 atoms, d = 8
    ncfg, nat, d = 12, 5, 8
    B_ = np.stack([np.add.reduce(rng.standard_normal((nat, d)), axis=0) for _ in range(ncfg)])
    dB = rng.standard_normal((ncfg * 3 * nat, d))
    e = rng.random(ncfg)
    f = rng.standard_normal(ncfg * 3 * nat)
    A = np.vstack([B_, dB])
    Ai = np.hstack([np.concatenate([np.ones(ncfg), np.zeros(dB.shape[0])])[:, None], A])
    b = np.concatenate([e, f])
    q = np.concatenate([100.0 * np.ones(ncfg), np.ones(dB.shape[0])])
    out["ne_A"], out["ne_b"], out["ne_q"] = A, b, q
    out["ne_M"], out["ne_v"] = truth_normal_eq(A, b, q, 0.0)
    out["ne_Mi"], out["ne_vi"] = truth_normal_eq(Ai, b, q, 0.01)
    # PCAState-like rows with a large mean
    Xc = 10.0 * rng.standard_normal(8) + rng.standard_normal((64, 8))
    out["cov_X"] = Xc
    out["cov_mean"], out["cov_centered"] = truth_cov(Xc, True)
    _, out["cov_raw"] = truth_cov(Xc, False)
    # ---- widened rows, appended after the draws above so the earlier vectors stay bit-identical ----
    # f2: GlobalSum / GlobalMean of ragged configurations, one of them beyond Julia's pairwise block size (1024 atoms)
    nat = [1, 2, 17, 96, 1500]
    mu6 = 3.0 * rng.standard_normal(6)
    loc = [mu6 + rng.standard_normal((n, 6)) for n in nat]
    out["gf_offsets"] = np.concatenate([[0], np.cumsum(nat)]).astype(np.int64)
    out["gf_local"] = np.concatenate(loc, axis=0)
    out["gf_sum"] = np.stack([l.astype(LD).sum(axis=0) for l in loc]).astype(np.float64)
    out["gf_mean"] = np.stack([l.astype(LD).sum(axis=0) / LD(len(l)) for l in loc]).astype(np.float64)
    # f4: InverseMultiquadric(Euclidean) (kernels.jl:156-164) on the config-1 slice
    Xl = f1.astype(LD)
    D2 = ((Xl[:, None, :] - Xl[None, :, :]) ** 2).sum(axis=2)
    out["c1_imq"] = ((LD(1.3) + D2 / LD(0.9) ** 2) ** LD(-0.5)).astype(np.float64)
    # f3: predictions A·β + β0 on the energy rows (Data/utils.jl:42-65)
    beta = rng.standard_normal(d)
    out["pr_beta"] = beta
    out["pr_y"] = (A.astype(LD) @ beta.astype(LD) + LD(0.37) * np.concatenate([np.ones(ncfg), np.zeros(dB.shape[0])]).astype(LD)).astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "hotpath_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "hotpath_golden.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
