"""plb200 — B200-native descriptor linear algebra for PotentialLearning.jl's hot path (host mirror of the Julia API).

Layout: csrc/ (sm_100a CUDA + C ABI -> plb200/libplb200.so), plb200/ (this package: ctypes binding + the reference's
operator interface for the path), julia/ (the ccall shim a Julia deployment loads). See DESIGN.md / INTEGRATION.md.
"""
from . import _lib
from ._lib import PLB200Error, finalize, init, last_timing, launch_count, pinned_zeros, workspace_bytes
from .data import (Configuration, DataSet, Energy, ForceDescriptors, Forces, LocalDescriptors, get_energy, get_force_descriptors,
                   get_forces, get_local_descriptors, get_values)
from .dimred import (PCA, ActiveSubspace, DimensionReducer, PCAState, compute_eigen, covariance, fit, fit_, project, select_eigendirections,
                     transform_)
from .dpp import (DBSCANSelector, EllEnsemble, RandomSelector, SubsetSelector, distance_matrix_kabsch, distance_matrix_periodic, get_dpp_mode, get_inclusion_prob, get_random_subset, greedy_subset, inclusion_prob, kDPP,
                  rescale_, sample, workspace_trim)
from .kernels import (RBF, CorrelationMatrix, Diagonal, Distance, Divergence, DotProduct, KernelSteinDiscrepancy, compute_divergence, Euclidean, Feature, GlobalMean, GlobalSum, InverseMultiquadric, Kernel,
                      KernelMatrix, Symmetric,
                      compute_distance, compute_feature, compute_features, compute_kernel, correlation_features, get_parameters, global_features, global_features_dev,
                      gram_tile_coords, kernel_matrix_dev, stack_locals)
from .learning import (CovariateLinearProblem, LinearBasisPotential, NormalEquationAccumulator, batch_sums, calc_all_metrics, get_metrics, mae,
                       mean_cos, rmse, rsq, LinearProblem, UnivariateLinearProblem, assemble_matrices,
                       get_all_energies, get_all_forces, learn_, learn_potential_, normal_equations, normal_equations_blocks, ols_sums, ooc_learn_, pack_rhs,
                       pack_rows, predict)

__all__ = [n for n in dir() if not n.startswith("_")]
