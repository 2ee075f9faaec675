#!/usr/bin/env python
"""The workflow of the reference's examples on the B200 path, with synthetic descriptors standing in for ACE
(examples/DPP-ACE-Si/fit-dpp-ace-si.jl + subsampling_utils.jl, examples/PCA-ACE-aHfO2/fit-pca-ace-ahfo2.jl):

  1. kDPP(ds, GlobalMean(), DotProduct())  -> L-ensemble (features, kernel matrix and eigendecomposition on the GPU)
  2. get_random_subset / get_inclusion_prob / get_dpp_mode
  3. learn!(lb, ds_train[subset], [100, 1], true)  -> weighted normal equations on the GPU, solve on the host
  4. get_all_energies / get_all_forces(ds_test, lb) -> predictions on the GPU, error metrics
  5. fit!(ds, PCAState) / transform!(ds, pca)       -> centred covariance, eigen and projections on the GPU

Run:  python examples/dpp_subsample_fit_pca.py [n_configs] [n_atoms] [n_descriptors]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "potentiallearning.jl_b200"))
import numpy as np

import plb200 as pl


def synthetic_dataset(rng, n, natoms, K):
    """Configurations whose energies / forces ARE linear in the descriptors (plus noise), like a converged ACE fit."""
    beta_true = rng.standard_normal(K) / np.sqrt(K)
    mu = 2.0 * rng.standard_normal(K)
    cfgs = []
    for _ in range(n):
        local = mu / natoms + rng.standard_normal((natoms, K)) / np.sqrt(natoms)
        fd = rng.standard_normal((natoms, 3, K))
        e = float(local.sum(0) @ beta_true + 0.37 + 1e-3 * rng.standard_normal())
        f = fd.reshape(-1, K) @ beta_true + 1e-3 * rng.standard_normal(3 * natoms)
        cfgs.append(pl.Configuration(pl.LocalDescriptors(local), pl.Energy(e), pl.ForceDescriptors(fd), pl.Forces(f.reshape(natoms, 3))))
    return pl.DataSet(cfgs), beta_true


def main(n=600, natoms=16, K=26, batch=60, verbose=True):
    say = print if verbose else (lambda *a, **k: None)
    rng = np.random.default_rng(0)
    pl.init(0)
    ds, beta_true = synthetic_dataset(rng, n, natoms, K)
    train, test = ds[0:n * 4 // 5], ds[n * 4 // 5:n]

    t0 = time.perf_counter()
    dpp = pl.kDPP(train, pl.GlobalMean(), pl.DotProduct(), batch_size=batch)           # dpp.jl:18-29
    subset = pl.get_random_subset(dpp, rng=np.random.default_rng(1))                    # dpp.jl:50-53
    incl = pl.get_inclusion_prob(dpp)                                                   # dpp.jl:68-70
    say(f"kDPP over {len(train)} configurations: batch {len(subset)}, sum of inclusion probabilities {incl.sum():.6f} "
        f"({(time.perf_counter() - t0) * 1e3:.1f} ms)")

    lb = pl.LinearBasisPotential(K)
    t0 = time.perf_counter()
    pl.learn_potential_(lb, train[subset], [100.0, 1.0], True)                          # learn!(lb, ds, ws, int)
    say(f"learn! on the DPP subset: |beta - beta_true| = {np.linalg.norm(lb.β - beta_true):.3e}, beta0 = {lb.β0[0]:.4f} "
        f"({(time.perf_counter() - t0) * 1e3:.1f} ms)")

    e_ref = np.array(pl.get_all_energies(test))
    f_ref = pl.get_all_forces(test)
    e_pred = pl.get_all_energies(test, lb)                                              # Data/utils.jl:42-49
    f_pred = pl.get_all_forces(test, lb)                                                # Data/utils.jl:58-65
    e_mae, f_mae = float(np.mean(np.abs(e_pred - e_ref))), float(np.mean(np.abs(f_pred - f_ref)))
    say(f"test set: energy MAE {e_mae:.3e}, force MAE {f_mae:.3e}")

    pca = pl.PCAState(tol=8)
    pl.fit_(ds, pca)                                                                    # pca_state.jl:20-40
    var_kept = float(np.sum(pca.λ[:8]) / np.sum(pca.λ))
    pl.transform_(ds, pca)                                                              # pca_state.jl:42-62
    new_dim = pl.get_values(pl.get_local_descriptors(ds[0])).shape[1]
    say(f"PCAState: kept 8 of {K} directions ({100 * var_kept:.1f} % of the variance); local descriptors are now {new_dim}-dimensional")
    say(f"plb200 kernels launched: {pl.launch_count()}")
    return {"subset": subset, "inclusion_sum": float(incl.sum()), "beta_err": float(np.linalg.norm(lb.β - beta_true)), "e_mae": e_mae,
            "f_mae": f_mae, "pca_dim": new_dim, "var_kept": var_kept}


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:4]]
    main(*(a + [600, 16, 26][len(a):]))
