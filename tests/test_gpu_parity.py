"""GPU parity tests: the CUDA path (through the C ABI / host mirror) against the CPU oracle on identical seeded inputs.

Tolerance: north_star asks for 1e-10 relative in fp64; the tests hold the kernels to RTOL below (tighter, so that
regressions show long before the contract is at risk). Kernel matrices are compared element-wise relative (all entries
are > 0), SYRK-type results norm-wise (max |diff| / max |ref|) plus element-wise where |ref| > 1e-3 max.
The first blocks re-state the reference's own tests (shapes d=8, 20 atoms, 10/8/100 configs).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-11  # contract: 1e-10


@pytest.fixture(scope="module")
def pl():
    import plb200

    plb200.init(0)
    return plb200


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle

    return oracle


def relerr(K, Kref):
    return float(np.max(np.abs(K - Kref) / np.abs(Kref)))


def normerr(C, Cref):
    scale = np.max(np.abs(Cref))
    e = float(np.max(np.abs(C - Cref)) / scale)
    big = np.abs(Cref) > 1e-3 * scale
    if big.any():
        e = max(e, float(np.max(np.abs(C - Cref)[big] / np.abs(Cref)[big])))
    return e


def fake_dataset(pl, rng, num_configs, num_atoms, d, energies=False, forces=False):
    cfgs = []
    for _ in range(num_configs):
        items = [pl.LocalDescriptors(rng.standard_normal((num_atoms, d)))]
        if energies:
            items.append(pl.Energy(rng.random()))
        if forces:
            items.append(pl.ForceDescriptors(rng.standard_normal((num_atoms, 3, d))))
            items.append(pl.Forces(rng.standard_normal((num_atoms, 3))))
        cfgs.append(pl.Configuration(*items))
    return pl.DataSet(cfgs)


# ---------------------------------------------------------------------------------------------------------------
# test/kernels/kernel_tests.jl re-stated
# ---------------------------------------------------------------------------------------------------------------
def test_reference_kernel_tests(pl, orc):
    rng = np.random.default_rng(7)
    d, num_atoms, num_configs = 8, 20, 10  # kernel_tests.jl:7-9
    ds = fake_dataset(pl, rng, num_configs, num_atoms, d)
    gm = pl.GlobalMean()
    f_gm = [pl.compute_feature(pl.get_local_descriptors(c), gm) for c in ds]
    assert f_gm[0].dtype == np.float64 and f_gm[0].shape == (d,)  # :30
    assert all(np.array_equal(a, b) for a, b in zip(pl.compute_features(ds, gm), f_gm))  # :32
    for c, f in zip(ds, f_gm):  # oracle: mean(get_values(B)) features.jl:33
        assert np.array_equal(f, orc.global_mean(list(pl.get_values(pl.get_local_descriptors(c)))))
    e = pl.Euclidean(d)
    assert pl.compute_distance(f_gm[0], f_gm[0], e) < np.finfo(float).eps  # :43
    assert pl.compute_distance(f_gm[0], f_gm[1], e) > 0.0  # :44
    dp, rbf_e = pl.DotProduct(), pl.RBF(e)
    assert np.isclose(pl.compute_kernel(f_gm[0], f_gm[0], dp), 1.0)  # :66
    assert np.isclose(pl.compute_kernel(f_gm[0], f_gm[0], rbf_e), 1.0)  # :67
    assert pl.compute_kernel(f_gm[0], f_gm[1], dp) > 0  # :72
    assert pl.compute_kernel(f_gm[0], f_gm[1], rbf_e) > 0  # :73
    Kd = pl.KernelMatrix(f_gm, dp)
    Kr = pl.KernelMatrix(f_gm, rbf_e)
    assert isinstance(Kd, pl.Symmetric) and Kd.data.dtype == np.float64  # :85
    assert isinstance(Kr, pl.Symmetric) and Kr.data.dtype == np.float64  # :87
    # value-level parity against the restated pair loop
    assert relerr(np.asarray(Kd), orc.symmetric(orc.kernel_matrix(f_gm, orc.DotProduct()))) < RTOL
    assert relerr(np.asarray(Kr), orc.symmetric(orc.kernel_matrix(f_gm, orc.RBF(orc.Euclidean(d))))) < RTOL
    # dataset forms (kernels.jl:250-269)
    Kds = pl.KernelMatrix(ds, gm, dp)
    assert np.array_equal(np.asarray(Kds), np.asarray(Kd))
    Kx = pl.KernelMatrix(ds[0:4], ds[4:10], gm, rbf_e)
    assert Kx.shape == (4, 6)
    assert relerr(Kx, orc.kernel_matrix_cross(f_gm[:4], f_gm[4:], orc.RBF(orc.Euclidean(d)))) < RTOL


def test_symmetric_storage_matches_reference(pl, orc):
    """Symmetric(K) over zeros(n,n): only the :U triangle of the storage is written (kernels.jl:216-222)."""
    rng = np.random.default_rng(8)
    for n in (5, 130, 300):
        X = rng.standard_normal((n, 12)) + 1.0
        S = pl.KernelMatrix(X, pl.DotProduct(), fill="U")
        assert np.array_equal(np.tril(S.data, -1), np.zeros((n, n)))
        ref = orc.kernel_matrix(X, orc.DotProduct())
        assert relerr(np.triu(S.data) + 1.0, np.triu(ref) + 1.0) < RTOL
        full = np.asarray(pl.KernelMatrix(X, pl.DotProduct()))
        assert np.array_equal(full, full.T), "exact symmetry"
        assert np.array_equal(np.triu(full), np.triu(S.data))


def test_symmetric_storage_streamed_path(pl, orc):
    """n >= 2048 with the :U storage contract takes the streamed host path (tile-column ranges + overlapped D2H)."""
    rng = np.random.default_rng(18)
    n, d = 2100, 20
    X = rng.standard_normal((n, d)) * 0.3 + 1.0
    for k, ko in ((pl.RBF(pl.Euclidean(d)), orc.RBF(orc.Euclidean(d))), (pl.DotProduct(), orc.DotProduct())):
        S = pl.KernelMatrix(X, k, fill="U")
        assert np.array_equal(np.tril(S.data, -1), np.zeros((n, n)))
        ref = orc.kernel_matrix(X, ko)
        iu = np.triu_indices(n)
        assert relerr(S.data[iu], ref[iu]) < RTOL
        full = np.asarray(pl.KernelMatrix(X, k))
        assert np.array_equal(np.triu(full), S.data), "streamed and one-shot paths agree bit for bit"


# ---------------------------------------------------------------------------------------------------------------
# Gram kernels: shapes, ragged edges, parameters
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d", [(1, 1), (2, 3), (7, 8), (127, 15), (128, 16), (129, 17), (300, 26), (513, 33), (1000, 26)])
def test_gram_dot_and_rbf_shapes(pl, orc, n, d):
    rng = np.random.default_rng(100 + n + d)
    mu = 2.0 * rng.standard_normal(d)
    X = mu + rng.standard_normal((n, d)) / np.sqrt(32)  # GlobalMean of 32 atoms of N(mu, 1)  (BASELINE config 1 generator)
    for alpha in (1, 2, 3):
        K = np.asarray(pl.KernelMatrix(X, pl.DotProduct(alpha)))
        ref = orc.symmetric(orc.kernel_matrix(X, orc.DotProduct(alpha)))
        assert relerr(K, ref) < RTOL, (n, d, alpha)
        assert np.array_equal(K, K.T)
        assert np.all(np.diag(K) == 1.0)
    K = np.asarray(pl.KernelMatrix(X, pl.RBF(pl.Euclidean(d), ell=0.7, beta=2.5)))
    ref = orc.symmetric(orc.kernel_matrix(X, orc.RBF(orc.Euclidean(d), ell=0.7, beta=2.5)))
    assert relerr(K, ref) < RTOL, (n, d)
    assert np.array_equal(K, K.T)
    assert np.all(np.diag(K) == 2.5), "K[i,i] == beta exactly (RBF.alpha is never applied, kernels.jl:73-74)"


@pytest.mark.parametrize("m,n,d", [(1, 1, 5), (3, 200, 8), (130, 257, 31), (300, 64, 26)])
def test_gram_cross(pl, orc, m, n, d):
    rng = np.random.default_rng(m * 1000 + n)
    X1 = rng.standard_normal((m, d)) * 0.5 + 1.0
    X2 = rng.standard_normal((n, d)) * 0.5 + 1.0
    for k, ko in ((pl.DotProduct(), orc.DotProduct()), (pl.RBF(pl.Euclidean(d)), orc.RBF(orc.Euclidean(d)))):
        K = pl.KernelMatrix(X1, X2, k)
        assert K.shape == (m, n)
        assert relerr(K, orc.kernel_matrix_cross(X1, X2, ko)) < RTOL


@pytest.mark.parametrize("d", [6, 7, 40])
def test_rbf_mahalanobis(pl, orc, d):
    """Euclidean(Cinv) with Diagonal and full Symmetric Cinv (distances.jl:46-51; the DPP-ACE-aHfO2-2 example uses a 716x716 Symmetric)."""
    rng = np.random.default_rng(d)
    n = 150
    X = rng.standard_normal((n, d)) * 0.4 + rng.standard_normal(d)
    cd = rng.uniform(0.5, 2.0, size=d)
    K = np.asarray(pl.KernelMatrix(X, pl.RBF(pl.Euclidean(pl.Diagonal(cd)), ell=1.3)))
    ref = orc.symmetric(orc.kernel_matrix(X, orc.RBF(orc.Euclidean(d, cd), ell=1.3)))
    assert relerr(K, ref) < RTOL
    B = rng.standard_normal((d, d))
    C = B @ B.T / d + np.eye(d)
    K = np.asarray(pl.KernelMatrix(X, pl.RBF(pl.Euclidean(pl.Symmetric(C)), ell=2.0)))
    ref = orc.symmetric(orc.kernel_matrix(X, orc.RBF(orc.Euclidean(d, C), ell=2.0)))
    assert relerr(K, ref) < RTOL
    assert np.array_equal(K, K.T)
    Kx = pl.KernelMatrix(X[:20], X[20:75], pl.RBF(pl.Euclidean(pl.Symmetric(C)), ell=2.0))
    assert relerr(Kx, orc.kernel_matrix_cross(X[:20], X[20:75], orc.RBF(orc.Euclidean(d, C), ell=2.0))) < RTOL


def test_rbf_full_cinv_at_example_scale(pl, orc):
    """A d x d Symmetric Cinv at the size the DPP-ACE-aHfO2-2 example uses (716 x 716 there; 301 here: odd, so the padded-copy
    path of the TMA operands is exercised too)."""
    rng = np.random.default_rng(77)
    n, d = 400, 301
    X = rng.standard_normal((n, d)) * 0.2 + rng.standard_normal(d)
    B = rng.standard_normal((d, d))
    C = B @ B.T / d + 0.5 * np.eye(d)
    K = np.asarray(pl.KernelMatrix(X, pl.RBF(pl.Euclidean(pl.Symmetric(C)), ell=6.0)))
    ref = orc.symmetric(orc.kernel_matrix(X, orc.RBF(orc.Euclidean(d, C), ell=6.0)))
    assert relerr(K, ref) < RTOL and np.array_equal(K, K.T)
    assert 1e-6 < np.median(ref) < 0.99


def test_inverse_multiquadric_and_correlation_features(pl, orc):
    """Widened rows on the same Gram engine: InverseMultiquadric(Euclidean) (kernels.jl:139-164) and
    CorrelationMatrix features (features.jl:37-50) with DotProduct / RBF / IMQ, as test/kernels/kernel_tests.jl exercises them."""
    rng = np.random.default_rng(55)
    n, d = 260, 11
    X = rng.standard_normal((n, d)) * 0.5 + 1.0
    for c2, ell in ((1.0, 1.0), (0.3, 2.5)):
        K = np.asarray(pl.KernelMatrix(X, pl.InverseMultiquadric(pl.Euclidean(d), c2=c2, ell=ell)))
        ref = orc.symmetric(orc.kernel_matrix(X, orc.InverseMultiquadric(orc.Euclidean(d), c2=c2, ell=ell)))
        assert relerr(K, ref) < RTOL and np.array_equal(K, K.T)
        assert np.allclose(np.diag(K), c2 ** -0.5, rtol=1e-15)
    cd = rng.uniform(0.5, 2.0, size=d)
    Kx = pl.KernelMatrix(X[:30], X[30:90], pl.InverseMultiquadric(pl.Euclidean(pl.Diagonal(cd)), c2=2.0))
    assert relerr(Kx, orc.kernel_matrix_cross(X[:30], X[30:90], orc.InverseMultiquadric(orc.Euclidean(d, cd), c2=2.0))) < RTOL
    with pytest.raises(AssertionError):
        pl.InverseMultiquadric(pl.Euclidean(d), c2=0.0)
    assert pl.get_parameters(pl.InverseMultiquadric(pl.Euclidean(d), c2=2.0, ell=3.0)) == (2.0, 3.0)
    # CorrelationMatrix features: 10 configs x 20 atoms x d = 8 (kernel_tests.jl:7-13, 27-33)
    ds = fake_dataset(pl, rng, 10, 20, 8)
    cm = pl.CorrelationMatrix()
    f_cm = pl.compute_features(ds, cm)
    assert isinstance(f_cm[0], pl.Symmetric) and f_cm[0].shape == (8, 8)  # :31
    ref_cm = [orc.correlation_matrix(list(pl.get_values(pl.get_local_descriptors(c)))) for c in ds]
    assert all(np.array_equal(np.asarray(a), b) for a, b in zip(f_cm, ref_cm))
    e8 = pl.Euclidean(8)
    assert np.isclose(pl.compute_kernel(f_cm[0], f_cm[0], pl.DotProduct()), 1.0)  # :68
    assert np.isclose(pl.compute_kernel(f_cm[0], f_cm[0], pl.RBF(e8)), 1.0)  # :69
    assert pl.compute_kernel(f_cm[0], f_cm[1], pl.DotProduct()) > 0 and pl.compute_kernel(f_cm[0], f_cm[1], pl.RBF(e8)) > 0  # :74-75
    for k, ko in ((pl.DotProduct(), orc.DotProduct()), (pl.RBF(e8), orc.RBF(orc.Euclidean(8))),
                  (pl.InverseMultiquadric(e8), orc.InverseMultiquadric(orc.Euclidean(8)))):
        K = pl.KernelMatrix(f_cm, k)
        assert isinstance(K, pl.Symmetric)  # :86, :88
        assert relerr(np.asarray(K), orc.symmetric(orc.kernel_matrix(ref_cm, ko))) < RTOL
    assert isinstance(pl.KernelMatrix(ds, cm, pl.DotProduct()), pl.Symmetric)
    with pytest.raises(TypeError):
        pl.KernelMatrix(f_cm, pl.RBF(pl.Euclidean(pl.Diagonal(np.ones(8)))))


def test_rbf_extreme_arguments(pl, orc):
    """exp argument outside the fast range (< -708: results below 3e-308) and the whole dynamic range in between."""
    rng = np.random.default_rng(9)
    n, d = 260, 10
    X = rng.standard_normal((n, d))
    for ell in (0.02, 0.05, 0.2, 5.0):
        K = np.asarray(pl.KernelMatrix(X, pl.RBF(pl.Euclidean(d), ell=ell)))
        ref = orc.symmetric(orc.kernel_matrix(X, orc.RBF(orc.Euclidean(d), ell=ell)))
        big = ref > 1e-290
        assert relerr(K[big], ref[big]) < 1e-10  # |x| up to ~700: the relative error of exp grows with |x| * eps
        assert np.all(np.abs(K[~big] - ref[~big]) <= 1e-290)
        assert np.all(np.isfinite(K))


def test_rbf_scale_sign_size_and_nan(pl, orc):
    """RBF.beta is any Real in the reference (kernels.jl:51-75): beta > 1 (the exponent of the fused 2^(y/256) is positive near the
    diagonal), beta = 1e6 / 1e-6, beta <= 0 (cannot be folded into the exponent: multiplied in afterwards), at depths whose two
    augmented columns fit the padding (d = 9, 26), start a new k-step (d = 15, 16, 32) or sit at its end (d = 14, 30). NaN inputs
    propagate to their row and column only."""
    rng = np.random.default_rng(31)
    for d in (9, 14, 15, 16, 26, 30, 32):
        n = 140
        X = rng.standard_normal((n, d)) * 0.5 + rng.standard_normal(d)
        for beta in (3.0, 1e6, 1e-6, -1.5, 0.0):
            K = np.asarray(pl.KernelMatrix(X, pl.RBF(pl.Euclidean(d), ell=1.1, beta=beta)))
            ref = orc.symmetric(orc.kernel_matrix(X, orc.RBF(orc.Euclidean(d), ell=1.1, beta=beta), use_c=False))
            if beta == 0.0:
                assert np.all(K == 0.0) and np.all(ref == 0.0)
            else:
                assert relerr(K, ref) < RTOL, (d, beta)
            assert np.array_equal(K, K.T) and np.all(np.diag(K) == beta)
        Kx = pl.KernelMatrix(X[:50], X[50:], pl.RBF(pl.Euclidean(d), ell=0.9, beta=-2.0))
        assert relerr(Kx, orc.kernel_matrix_cross(X[:50], X[50:], orc.RBF(orc.Euclidean(d), ell=0.9, beta=-2.0))) < RTOL
    Xn = rng.standard_normal((130, 9))
    Xn[5, 3] = np.nan
    Kn = np.asarray(pl.KernelMatrix(Xn, pl.RBF(pl.Euclidean(9))))
    assert np.isnan(np.delete(Kn[5], 5)).all() and np.isnan(np.delete(Kn[:, 5], 5)).all()
    assert not np.isnan(np.delete(np.delete(Kn, 5, 0), 5, 1)).any()


def test_gram_mean_heavy_descriptors_keep_1e10(pl, orc):
    """|x|^2 >> d2 (GlobalMean descriptors with a large common mean): the column-mean shift keeps the Gram-form distance exact."""
    rng = np.random.default_rng(5)
    n, d = 400, 300
    X = 2.0 * rng.standard_normal(d) + rng.standard_normal((n, d)) / np.sqrt(96)
    K = np.asarray(pl.KernelMatrix(X, pl.RBF(pl.Euclidean(d))))
    ref = orc.symmetric(orc.kernel_matrix(X, orc.RBF(orc.Euclidean(d))))
    assert relerr(K, ref) < 1e-12


# ---------------------------------------------------------------------------------------------------------------
# test/learning/linear_tests.jl re-stated + normal-equation parity
# ---------------------------------------------------------------------------------------------------------------
def test_reference_univariate_linear_tests(pl, orc):
    rng = np.random.default_rng(0)  # linear_tests.jl:5 seeds the RNG
    d, num_atoms, num_configs = 8, 20, 8  # :20-22
    ds = fake_dataset(pl, rng, num_configs, num_atoms, d, energies=True)
    lp = pl.LinearProblem(ds)
    assert isinstance(lp, pl.LinearProblem) and isinstance(lp, pl.UnivariateLinearProblem)  # :30-31
    assert lp.dv_data == [pl.get_values(pl.get_energy(c)) for c in ds]  # :36
    sums = [np.add.reduce(pl.get_values(pl.get_local_descriptors(c)), axis=0) for c in ds]
    assert all(np.array_equal(a, b) for a, b in zip(lp.iv_data, sums))  # :37
    assert np.array_equal(lp.σ, [1.0]) and np.array_equal(lp.β, np.zeros(d)) and np.array_equal(lp.β0, np.zeros(1))  # :38-40
    A = np.stack(sums)
    e = np.array(lp.dv_data)
    pl.learn_(lp)  # default: ws = [1.0], int = false (:44, :278-284)
    lp2 = pl.LinearProblem(ds)
    pl.learn_(lp2, [1.0], False)
    assert np.array_equal(lp.β, lp2.β) and np.array_equal(lp.β0, lp2.β0)  # :48-49
    assert np.max(np.abs(A @ lp2.β + lp2.β0 - e)) < 1e-4  # :50 (8 equations, 8 unknowns)
    # 8 equations in 9 unknowns is singular with an intercept (the reference hits its pinv fallback, :199-205):
    # use 12 configurations for the parity part of the intercept case
    ds12 = fake_dataset(pl, rng, 12, num_atoms, d, energies=True)
    lp3 = pl.LinearProblem(ds12)
    pl.learn_(lp3, [1.0], True)  # :52-54
    o3 = orc.linear_problem([list(pl.get_values(pl.get_local_descriptors(c))) for c in ds12], [pl.get_values(pl.get_energy(c)) for c in ds12])
    orc.learn_wls(o3, [1.0], True)
    assert np.allclose(lp3.β, o3.beta, rtol=1e-8, atol=1e-10) and np.allclose(lp3.β0, o3.beta0, rtol=1e-8, atol=1e-10)
    # OLS (:60-64): AtA / Atb sums on the GPU vs sum(v*v') restated
    AtA, Atb = pl.ols_sums(lp.iv_data, lp.dv_data)
    AtA_ref, Atb_ref = orc.ols_sums(lp.iv_data, lp.dv_data)
    assert normerr(AtA, AtA_ref) < RTOL and normerr(Atb, Atb_ref) < RTOL
    lp4 = pl.LinearProblem(ds)
    pl.learn_(lp4, 1e-8)
    assert np.max(np.abs(A @ lp4.β - e)) < 1e-4  # :64


def test_ols_pinv_cutoff_is_relative_like_julia(pl, orc):
    """learn!(lp, α) uses pinv(AtA, α) (linear-learn.jl:19) — in Julia that is rtol = α: the cutoff is α * σ_max, not α. A spectrum that
    straddles both readings (σ = 400, 25, 1, 0.04 ... with α = 0.01: the relative cutoff 4 drops what an absolute 0.01 would keep)."""
    rng = np.random.default_rng(77)
    K, n = 6, 40
    Qm, _ = np.linalg.qr(rng.standard_normal((K, K)))
    scales = np.array([20.0, 5.0, 1.0, 0.2, 0.05, 0.01])  # AtA singular values ~ n * scales^2
    iv = [Qm @ (scales * rng.standard_normal(K)) for _ in range(n)]
    dv = [float(rng.standard_normal()) for _ in range(n)]
    alpha = 0.01
    AtA, _ = orc.ols_sums(iv, dv)
    s = np.linalg.svd(AtA, compute_uv=False)
    assert (s > alpha).sum() > (s > alpha * s.max()).sum() >= 1, "the two cutoffs must select different ranks for this test to mean anything"
    lp = pl.UnivariateLinearProblem(iv, dv, np.zeros(K), np.zeros(1), np.ones(1), np.zeros((K, K)))
    pl.learn_(lp, alpha)
    beta, sigma, Sigma = orc.learn_ols(iv, dv, alpha)
    assert np.allclose(lp.β, beta, rtol=1e-9, atol=1e-12 * np.abs(beta).max())
    assert abs(lp.σ[0] - sigma) < 1e-9 * sigma and np.allclose(lp.Σ, Sigma, rtol=1e-8, atol=1e-10 * np.abs(Sigma).max())
    import scipy.linalg

    assert np.allclose(pl.learning._pinv_rtol(AtA, alpha), scipy.linalg.pinv(AtA, atol=0.0, rtol=alpha), rtol=1e-9, atol=1e-14)


def test_blocks_orientation_follows_shape(pl, orc):
    """entries of lp.dB are K x m whatever their memory order (Fortran like Julia's, C after a numpy copy), entries of lp.B K-vectors —
    including K == 1 and 3 natoms == K, where memory order says nothing."""
    rng = np.random.default_rng(78)
    for K, ms in ((5, (5, 3, 5)), (1, (4, 1, 2)), (7, (2, 9))):
        B = [rng.standard_normal(K) for _ in ms]
        dBf = [np.asfortranarray(rng.standard_normal((K, m))) for m in ms]
        dBc = [np.ascontiguousarray(x) for x in dBf]
        A = np.concatenate([np.stack(B)] + [x.T for x in dBf], axis=0)
        b = rng.standard_normal(A.shape[0])
        Mo, vo = orc.normal_equations(A, b, np.ones(A.shape[0]), 0.0)
        for dB in (dBf, dBc):
            M, v = pl.normal_equations_blocks(B + dB, b)
            assert normerr(M, Mo) < RTOL and normerr(v, vo) < RTOL
        M, v = pl.normal_equations_blocks([np.stack(B)] + [x.T.copy() for x in dBf], b, layout="rows")
        assert normerr(M, Mo) < RTOL and normerr(v, vo) < RTOL


def test_reference_covariate_linear_tests(pl, orc):
    rng = np.random.default_rng(1)
    d, num_atoms, num_configs = 8, 20, 100  # :71-73
    ds = fake_dataset(pl, rng, num_configs, num_atoms, d, energies=True, forces=True)
    lp = pl.LinearProblem(ds)
    assert isinstance(lp, pl.CovariateLinearProblem)  # :90
    # layout pinned by linear_tests.jl:93-97
    assert lp.e == [pl.get_values(pl.get_energy(c)) for c in ds]
    for c, f, B, dB in zip(ds, lp.f, lp.B, lp.dB):
        fv = pl.get_values(pl.get_forces(c))
        assert np.array_equal(f, np.concatenate(list(fv)))  # reduce(vcat, fi)
        assert np.array_equal(B, np.add.reduce(pl.get_values(pl.get_local_descriptors(c)), axis=0))
        fd = pl.get_values(pl.get_force_descriptors(c))
        cols = [fd[a, x] for a in range(num_atoms) for x in range(3)]  # reduce(vcat, get_values(fd)) then reduce(hcat, ·)
        assert dB.shape == (d, 3 * num_atoms) and np.array_equal(dB, np.stack(cols, axis=1))
    o = orc.linear_problem(
        [list(pl.get_values(pl.get_local_descriptors(c))) for c in ds], [pl.get_values(pl.get_energy(c)) for c in ds],
        [[list(a) for a in pl.get_values(pl.get_force_descriptors(c))] for c in ds], [list(pl.get_values(pl.get_forces(c))) for c in ds])
    for ws, intercept, lam in (([1.0, 1.0], False, 0.0), ([1.0, 1.0], True, 0.0), ([100.0, 1.0], True, 0.01), ([30.0, 1.0], False, 0.01)):
        lpi = pl.LinearProblem(ds)
        pl.learn_(lpi, ws, intercept, λ=lam)
        A, W, b = pl.assemble_matrices(lpi, ws)
        M, v = pl.normal_equations(A, b, None, len(lpi.e), ws[0], ws[1], intercept, lam)
        Ao, bo, qo = orc.wls_matrices(o, ws, intercept)
        Mo, vo = orc.normal_equations(Ao, bo, qo, lam)
        Mc, vc = orc.normal_equations(Ao, bo, qo, lam, use_c=True)  # reference operation order (Q*A then A'*(QA))
        assert normerr(M, Mo) < RTOL and normerr(v, vo) < RTOL
        assert normerr(M, Mc) < RTOL and normerr(v, vc) < RTOL
        assert np.array_equal(M, M.T), "exactly symmetric"
        orc.learn_wls(o, ws, intercept, lam)
        assert np.allclose(lpi.β, o.beta, rtol=1e-9, atol=1e-12)
        if intercept:  # without an intercept β0 is left untouched (linear-learn.jl:261-266)
            assert np.allclose(lpi.β0, o.beta0, rtol=1e-9, atol=1e-12)
        else:
            assert np.array_equal(lpi.β0, np.zeros(1))
    # default learn! == ws=[1,1], int=false (:104-109)
    lpa, lpb = pl.LinearProblem(ds), pl.LinearProblem(ds)
    pl.learn_(lpa)
    pl.learn_(lpb, [1.0, 1.0], False)
    assert np.array_equal(lpa.β, lpb.β) and np.array_equal(lpa.β0, lpb.β0)
    # MLE assembly blocks (:45-49)
    AtAe, Atbe, AtAf, Atbf = pl.learn_(pl.LinearProblem(ds), 1e-8)
    re, rbe = orc.ols_sums(o.B, o.e)
    rf, rbf = orc.ols_sums(o.dB, o.f)
    assert normerr(AtAe, re) < RTOL and normerr(Atbe, rbe) < RTOL and normerr(AtAf, rf) < RTOL and normerr(Atbf, rbf) < RTOL


# K = 127, 128, 255 exercise the weighted-column-sum fallback for the moment vectors, the others the augmented columns
@pytest.mark.parametrize("R,K,ne", [(1, 1, 1), (17, 3, 5), (1000, 127, 0), (4099, 128, 99), (20000, 500, 700), (3000, 300, 3000), (2500, 255, 40), (5000, 142, 5000),
                                    (70000, 14, 123)])
def test_normal_equations_shapes(pl, orc, R, K, ne):
    rng = np.random.default_rng(R + K)
    A = rng.standard_normal((R, K))
    b = rng.standard_normal(R)
    for intercept, lam, (we, wf) in ((False, 0.0, (1.0, 1.0)), (True, 0.01, (100.0, 1.0))):
        M, v = pl.normal_equations(A, b, None, ne, we, wf, intercept, lam)
        q = np.concatenate([we * np.ones(ne), wf * np.ones(R - ne)])
        Ai = np.hstack([np.concatenate([np.ones(ne), np.zeros(R - ne)])[:, None], A]) if intercept else A
        Mo, vo = orc.normal_equations(Ai, b, q, lam)
        assert normerr(M, Mo) < RTOL, (R, K, ne, intercept)
        assert normerr(v, vo) < RTOL
        assert np.array_equal(M, M.T)
    w = rng.uniform(0.1, 3.0, size=R)  # per-row weights (ooc_learn! energy weights, linear-learn.jl:341-352)
    M, v = pl.normal_equations(A, b, w)
    Mo, vo = orc.normal_equations(A, b, w, 0.0)
    assert normerr(M, Mo) < RTOL and normerr(v, vo) < RTOL


def test_ooc_learn_accumulation(pl, orc):
    rng = np.random.default_rng(3)
    K = 24
    configs = []
    for natoms in (4, 7, 4, 12, 9, 7):
        configs.append((rng.standard_normal(K) * natoms, rng.standard_normal((3 * natoms, K)), rng.standard_normal() * natoms, rng.standard_normal(3 * natoms)))
    for norm in (None, "standard", "squared"):
        beta = np.zeros(K)
        AtWA, AtWb = pl.ooc_learn_(beta, configs, ws=(30.0, 1.0), λ=0.01, eweight_normalized=norm)
        Ao, bo = orc.ooc_accumulate(configs, ws=(30.0, 1.0), eweight_normalized=norm)
        assert normerr(AtWA, Ao + 0.01 * np.eye(K)) < RTOL and normerr(AtWb, bo) < RTOL
        assert np.allclose(beta, np.linalg.solve(Ao + 0.01 * np.eye(K), bo), rtol=1e-9)


# ---------------------------------------------------------------------------------------------------------------
# test/dimension_reduction/dimension_reduction_tests.jl re-stated + covariance parity
# ---------------------------------------------------------------------------------------------------------------
def test_reference_dimension_reduction_tests(pl, orc):
    rng = np.random.default_rng(11)
    d, num_atoms, num_configs = 8, 20, 10
    ds = fake_dataset(pl, rng, num_configs, num_atoms, d)

    def gradQ(c):
        return np.add.reduce(pl.get_values(pl.get_local_descriptors(c)), axis=0)

    num_dim = 2
    as_ = pl.ActiveSubspace(lambda c: 0.5 * gradQ(c) @ gradQ(c), gradQ, num_dim)
    pca = pl.PCA(num_dim)
    lam_as, W_as = pl.fit(ds, as_)
    lam_pca, W_pca = pl.fit(ds, pca)
    assert lam_as.dtype == np.float64 and W_as.shape == (d, num_dim)  # :30-33
    assert lam_pca.dtype == np.float64 and W_pca.shape == (d, num_dim)  # :36-39
    assert np.allclose(lam_as, lam_pca)  # :41
    rows = [gradQ(c) for c in ds]
    lam_o, W_o = orc.select_eigendirections(rows, num_dim)
    assert np.allclose(lam_pca, lam_o, rtol=1e-10, atol=1e-12)
    # PCAState (untested in the reference; exercised by examples/PCA-ACE-aHfO2)
    st = pl.PCAState(tol=num_dim)
    pl.fit_(ds, st)
    so = orc.pcastate_fit([list(pl.get_values(pl.get_local_descriptors(c))) for c in ds], orc.PCAState(tol=num_dim))
    assert np.allclose(st.m, so.m, rtol=1e-13) and np.allclose(st.λ, so.lam, rtol=1e-9, atol=1e-12)
    m_first = st.m.copy()
    pl.fit_(ds[0:5], st)  # the mean persists across fits (pca_state.jl:34)
    assert np.array_equal(st.m, m_first)
    st_f = pl.PCAState(tol=0.05)
    pl.fit_(ds, st_f)
    lam_f, W_f = orc.select_eigendirections([r - so.m for r in rows], 0.05)
    assert st_f.W.shape == W_f.shape


# K = 15, 128, 255: shapes where the rho column cannot ride along as an augmented column (K % 16 == 15 or K % 128 == 0): the shifted
# column-sum pass closes the one-sided centring instead
@pytest.mark.parametrize("n,K", [(1, 1), (10, 8), (999, 129), (5000, 1000), (40000, 64), (700, 15), (3000, 128), (2000, 255)])
def test_covariance_parity(pl, orc, n, K):
    rng = np.random.default_rng(n + K)
    X = 10.0 * rng.standard_normal(K) + rng.standard_normal((n, K))  # mean/sigma ~ 10: kills one-pass formulas
    C, m = pl.covariance(X, center=True)
    mo, Co = orc.centered_cov(X)
    assert np.max(np.abs(m - mo) / np.abs(mo)) < 1e-13
    if n > 1:
        assert normerr(C, Co) < RTOL
    assert np.array_equal(C, C.T)
    C0, _ = pl.covariance(X, center=False)
    assert normerr(C0, (X.T @ X) / n) < RTOL
    if n * K * K <= 10 * 8 * 8:
        assert normerr(C0, orc.second_moment(X, use_c=False)) < RTOL  # sequential outer-product sum, reference order
    given = mo + 0.5
    C1, m1 = pl.covariance(X, mean=given, center=True)
    assert np.array_equal(m1, given)
    assert normerr(C1, ((X - given).T @ (X - given)) / n) < RTOL


# ---------------------------------------------------------------------------------------------------------------
# subset selection (test/subset_selector/subset_selector_tests.jl): types only, as in the reference
# ---------------------------------------------------------------------------------------------------------------
def test_reference_kdpp_tests(pl):
    rng = np.random.default_rng(21)
    ds = fake_dataset(pl, rng, 10, 20, 8)
    dpp = pl.kDPP(ds, pl.GlobalMean(), pl.DotProduct(), batch_size=2)  # :16
    assert isinstance(dpp, pl.SubsetSelector)
    sub = pl.get_random_subset(dpp, rng=np.random.default_rng(0))
    assert len(sub) == 2 and all(isinstance(i, int) for i in sub)
    mode = pl.get_dpp_mode(dpp)
    assert len(mode) == 2 and all(isinstance(i, int) for i in mode)
    ip = pl.get_inclusion_prob(dpp)
    assert ip.dtype == np.float64 and ip.shape == (10,)
    assert abs(ip.sum() - 2.0) < 1e-6  # expected cardinality after rescale!(ell, batch_size)


# ---------------------------------------------------------------------------------------------------------------
# device-resident API, tile sharding, full BASELINE sizes through size-independent properties
# ---------------------------------------------------------------------------------------------------------------
def test_device_api_tile_sharding(pl, orc):
    import torch

    rng = np.random.default_rng(31)
    n, d = 700, 40  # T = 6 tile rows: even T; 701 -> odd handled below
    for n in (700, 897):
        X = rng.standard_normal((n, d)) * 0.3 + 1.0
        Xd = torch.as_tensor(X, device="cuda")
        k = pl.RBF(pl.Euclidean(d))
        full = pl.kernel_matrix_dev(Xd, k).cpu().numpy()
        ref = orc.symmetric(orc.kernel_matrix(X, orc.RBF(orc.Euclidean(d))))
        assert relerr(full, ref) < RTOL
        nparts = 3
        seen = set()
        for part in range(nparts):
            tiles = pl.kernel_matrix_dev(Xd, k, part=part, nparts=nparts, packed=True).cpu().numpy()
            coords = pl.gram_tile_coords(n, 0, part, nparts)
            assert tiles.shape[0] == len(coords)
            for (ti, tj), tile in zip(coords, tiles):
                if ti < 0:
                    continue
                assert (ti, tj) not in seen
                seen.add((ti, tj))
                i0, j0 = 128 * ti, 128 * tj
                blk = full[i0:i0 + 128, j0:j0 + 128]
                assert np.array_equal(tile[: blk.shape[0], : blk.shape[1]], blk)
        T = (n + 127) // 128
        assert len(seen) == T * (T + 1) // 2
    torch.cuda.synchronize()


@pytest.mark.parametrize("n", [3000, 9000])
def test_rbf_medium_sizes_all_depth_classes(pl, n):
    """Both tile shapes of the RBF kernel (64 x 64 below ~8600 features, 128 x 128 above) at contraction depths whose two augmented
    columns fit the padding (d = 13, 30, 127 -> free), fill it exactly (d = 14) or open another k-step (d = 15, 16, 64), for the three
    kinds of Cinv, `Symmetric(:U)` storage and both triangles: 150 sampled rows against numpy fp64 at the contract's 1e-10,
    exact symmetry, exact diagonal."""
    import torch

    rng = np.random.default_rng(n)
    rows = np.sort(rng.choice(n, size=150, replace=False))
    for trial, d in enumerate((13, 14, 15, 16, 30, 64, 127)):
        X = rng.standard_normal((n, d)) * 0.35 + 1.5 * rng.standard_normal(d)
        ell, beta = 0.9 + 0.05 * trial, (0.5, 1.0, 3.0)[trial % 3]
        kind = trial % 3
        if kind == 0:
            e, C = pl.Euclidean(d), np.eye(d)
        elif kind == 1:
            cd = rng.uniform(0.5, 2.0, size=d)
            e, C = pl.Euclidean(pl.Diagonal(cd)), np.diag(cd)
        else:
            B = rng.standard_normal((d, d))
            C = B @ B.T / d + np.eye(d)
            e = pl.Euclidean(pl.Symmetric(C))
        Xd = torch.as_tensor(X, device="cuda")
        K = pl.kernel_matrix_dev(Xd, pl.RBF(e, ell=ell, beta=beta))
        torch.cuda.synchronize()
        assert bool(torch.equal(K, K.T)) and bool(torch.all(torch.diagonal(K) == beta)), (n, d)
        ref = np.empty((len(rows), n))
        for a, r in enumerate(rows):
            U = X[r] - X
            ref[a] = beta * np.exp(-0.5 * np.einsum("jk,jk->j", U @ C, U) / ell ** 2)
        got = K[torch.as_tensor(rows, device="cuda")].cpu().numpy()
        ok = ref > 1e-280
        assert np.max(np.abs(got[ok] - ref[ok]) / ref[ok]) < 1e-10, (n, d, kind)
        S = np.asarray(pl.KernelMatrix(X, pl.RBF(e, ell=ell, beta=beta))) if trial in (1, 5) else None  # host path, Symmetric(:U) storage
        if S is not None:
            assert np.array_equal(S[rows], got), "host path (streamed / staircase copy) and device-resident path agree bit for bit"


def test_config2_full_size_properties(pl, orc):
    """BASELINE config 2: RBF(Euclidean(300)), N = 20 000, d = 300 on one GPU — checked through size-independent properties:
    exact symmetry, K[i,i] == beta, sampled rows against the restated pair loop."""
    import torch

    n, d = 20000, 300
    g = torch.Generator(device="cuda").manual_seed(1002)
    mu = 2.0 * torch.randn(d, device="cuda", dtype=torch.float64, generator=g)
    X = mu + torch.randn((n, d), device="cuda", dtype=torch.float64, generator=g) / np.sqrt(96.0)
    K = pl.kernel_matrix_dev(X, pl.RBF(pl.Euclidean(d)))
    torch.cuda.synchronize()
    assert bool(torch.equal(K, K.T))
    assert bool(torch.all(torch.diagonal(K) == 1.0))
    rows = [0, 1, 127, 128, 9999, 19871, 19999]
    Xh = X.cpu().numpy()
    Kh = K[rows].cpu().numpy()
    ref = orc.kernel_matrix_cross(Xh[rows], Xh, orc.RBF(orc.Euclidean(d)))
    off = np.ones_like(ref, dtype=bool)
    for a, r in enumerate(rows):
        off[a, r] = False  # diagonal is exactly beta by contract; the pair loop gives exp(-0) = 1 too
    assert relerr(Kh[off], ref[off]) < RTOL
    assert 1e-3 < float(np.median(ref)) < 0.9, "kernel values are non-degenerate (SURVEY 8d generator)"
    del K


def test_config3_slice_properties(pl, orc):
    """BASELINE config 3 shape at 1/10 scale (1 000 configs x 96 atoms, K = 500): linearity in the row shards
    (sum of two shard results == whole) and parity on the whole."""
    import torch

    from plb200.device import DeviceEngine

    eng = DeviceEngine()
    ncfg, natoms, K = 1000, 96, 500
    ne, nf = ncfg, ncfg * 3 * natoms
    g = torch.Generator(device="cuda").manual_seed(1003)
    A = torch.randn((ne + nf, K), device="cuda", dtype=torch.float64, generator=g)
    A[:ne] *= np.sqrt(96.0)
    b = torch.randn(ne + nf, device="cuda", dtype=torch.float64, generator=g)
    S, vec = eng.syrk_partial(A, None, ne, 100.0, 1.0, b, None, True)
    half = ne + nf // 2
    S1, v1 = eng.syrk_partial(A[:half], None, ne, 100.0, 1.0, b[:half], None, True)
    S2, v2 = eng.syrk_partial(A[half:], None, 0, 100.0, 1.0, b[half:], None, True)
    torch.cuda.synchronize()
    assert float(((S1 + S2) - S).abs().max() / S.abs().max()) < 1e-13
    assert float(((v1 + v2) - vec).abs().max() / vec.abs().max()) < 1e-13
    M, v = eng.normal_eq_finish(S, vec, True, 0.01)
    Ah, bh = A.cpu().numpy(), b.cpu().numpy()
    q = np.concatenate([100.0 * np.ones(ne), np.ones(nf)])
    Ai = np.hstack([np.concatenate([np.ones(ne), np.zeros(nf)])[:, None], Ah])
    Mo, vo = orc.normal_equations(Ai, bh, q, 0.01)
    assert normerr(M.cpu().numpy(), Mo) < RTOL and normerr(v.cpu().numpy(), vo) < RTOL
    assert bool(torch.equal(M, M.T))


# ---------------------------------------------------------------------------------------------------------------
# the raw C ABI (what a Julia ccall sees): layouts, fill modes, sharded host entry, error codes
# ---------------------------------------------------------------------------------------------------------------
def test_c_abi_layouts_and_errors(pl, orc):
    import ctypes

    from plb200 import _lib

    lib = _lib.lib()
    rng = np.random.default_rng(41)
    n, d = 150, 9
    X = rng.standard_normal((n, d)) + 0.5
    ref = orc.symmetric(orc.kernel_matrix(X, orc.DotProduct()))
    # Julia layout: column-major d x n features with a padded leading dimension, K with ldk > n, uplo = :U only
    ldx, ldk = d + 3, n + 5
    Xj = np.zeros((n, ldx))
    Xj[:, :d] = X
    Kj = np.full((n, ldk), -7.0)  # sentinel: untouched entries must keep it
    _lib.check(lib.pl_gram_dot(_lib.dptr(Xj), n, d, ldx, 2, _lib.dptr(Kj), ldk, _lib.PL_FILL_JULIA_U), "pl_gram_dot")
    # Julia K[i,j] (i <= j) lives at K[i + j*ldk] = row-major element [j][i]
    got = Kj[:, :n]
    il = np.tril_indices(n)
    # zero-mean-ish features: some pairs are nearly orthogonal, where a.b cancels and ANY summation order (the reference's
    # BLAS ddot included) moves the tiny kernel value by ~eps * sum|a_k b_k|: relative + small absolute tolerance here
    assert np.allclose(got[il], ref[il], rtol=RTOL, atol=1e-14)
    assert np.all(Kj[:, n:] == -7.0), "padding columns of the caller's buffer are not written"
    iu = np.triu_indices(n, 1)
    assert np.all((got[iu] == 0.0) | (got[iu] == -7.0)), "other triangle: zeros inside diagonal tiles, untouched elsewhere"
    # rectangular: Julia column-major m x n == row-major n x m -> pass F2 first
    F1, F2 = X[:40], X[40:150]
    Kc = np.zeros((110, 40))
    _lib.check(lib.pl_gram_cross(_lib.PL_KERNEL_RBF, _lib.dptr(np.ascontiguousarray(F2)), 110, d, _lib.dptr(np.ascontiguousarray(F1)), 40, d, d, 0,
                                 0, None, 1.0, 1.0, _lib.dptr(Kc), 40), "pl_gram_cross")
    assert relerr(Kc.T, orc.kernel_matrix_cross(F1, F2, orc.RBF(orc.Euclidean(d)))) < RTOL
    # sharded host entry: union of the parts == the full matrix
    nparts = 3
    full = orc.symmetric(orc.kernel_matrix(X, orc.RBF(orc.Euclidean(d))))
    seen = 0
    for part in range(nparts):
        cnt = int(lib.pl_gram_tile_count(n, 0, part, nparts))
        tiles = np.zeros((cnt, 128, 128))
        _lib.check(lib.pl_gram_part(_lib.PL_KERNEL_RBF, _lib.dptr(np.ascontiguousarray(X)), n, d, d, 0, 0, None, 1.0, 1.0, part, nparts,
                                    _lib.dptr(tiles.reshape(-1))), "pl_gram_part")
        for (ti, tj), tile in zip(pl.gram_tile_coords(n, 0, part, nparts), tiles):
            if ti < 0:
                continue
            blk = full[128 * ti:128 * ti + 128, 128 * tj:128 * tj + 128]
            assert relerr(tile[: blk.shape[0], : blk.shape[1]], blk) < RTOL
            seen += 1
    assert seen == 3  # T = 2 -> 3 upper-triangle tiles
    # error behaviour: integer status + message, never an exception across the ABI, output untouched
    K0 = np.zeros((4, 4))
    assert lib.pl_gram_dot(_lib.dptr(X), 4, 0, 1, 2, _lib.dptr(K0), 4, 3) == 1 and b"bad" in lib.pl_last_error()  # PL_EINVAL
    assert lib.pl_gram_dot(_lib.dptr(X), 4, d, d - 1, 2, _lib.dptr(K0), 4, 3) == 1
    assert lib.pl_gram_dot(_lib.dptr(X), 4, d, d, 2, _lib.dptr(K0), 4, 7) == 1  # bad fill
    assert lib.pl_gram_rbf(_lib.dptr(X), 4, d, d, 1, None, 1.0, 1.0, _lib.dptr(K0), 4, 3) == 1  # diagonal Cinv without data
    assert lib.pl_gram_rbf(_lib.dptr(X), 4, d, d, 0, None, 0.0, 1.0, _lib.dptr(K0), 4, 3) == 1  # ell == 0
    assert lib.pl_normal_eq(_lib.dptr(X), n, d, d, None, 0, 1.0, 1.0, None, 0, 0.0, _lib.dptr(K0), _lib.dptr(np.zeros(d))) == 1  # AtWb without b
    assert np.all(K0 == 0.0)
    with pytest.raises(ValueError, match="DimensionMismatch"):
        pl.KernelMatrix(X, pl.RBF(pl.Euclidean(d + 1)))
    # empty input is a no-op, not an error
    assert lib.pl_gram_dot(_lib.dptr(X), 0, d, d, 2, _lib.dptr(K0), 4, 3) == 0
    t = pl.last_timing()
    assert set(t) == {"kernel_ms", "h2d_ms", "d2h_ms", "host_wall_ms"} and pl.launch_count() > 0


# ---------------------------------------------------------------------------------------------------------------
# after the fit: predictions and the PCAState projection
# ---------------------------------------------------------------------------------------------------------------
def test_predictions_and_transform(pl, orc):
    rng = np.random.default_rng(91)
    d, nat, ncfg = 8, 20, 30
    ds = fake_dataset(pl, rng, ncfg, nat, d, energies=True, forces=True)
    lb = pl.LinearBasisPotential(d)
    lp = pl.learn_potential_(lb, ds, [100.0, 1.0], True, λ=0.01)  # learn!(lb, ds, ws, int) (linear-learning-problem.jl:158-170)
    assert np.array_equal(lb.β, lp.β) and np.array_equal(lb.β0, lp.β0)  # linear_tests.jl:106-107
    e_pred = pl.get_all_energies(ds, lb)
    f_pred = pl.get_all_forces(ds, lb)
    B = np.stack(lp.B)
    dB = np.concatenate([m.T for m in lp.dB])
    assert np.allclose(e_pred, B @ lb.β + lb.β0[0], rtol=1e-13, atol=1e-13)
    assert np.allclose(f_pred, dB @ lb.β, rtol=1e-12, atol=1e-13)
    assert np.array_equal(pl.get_all_forces(ds), np.concatenate(lp.f)) and pl.get_all_energies(ds) == lp.e
    for R, K in ((1, 1), (1000, 7), (5000, 500)):
        A = rng.standard_normal((R, K))
        beta = rng.standard_normal(K)
        y = pl.predict(A, beta, 0.25, R // 3)
        ref = A @ beta + np.concatenate([0.25 * np.ones(R // 3), np.zeros(R - R // 3)])
        assert np.max(np.abs(y - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))
    # transform!(ds, pca) (pca_state.jl:42-62) against the formula evaluated on the host
    st = pl.PCAState(tol=3)
    pl.fit_(ds, st)
    ld0 = [pl.get_values(pl.get_local_descriptors(c)).copy() for c in ds]
    fd0 = [pl.get_values(pl.get_force_descriptors(c)).copy() for c in ds]
    pl.transform_(ds, st)
    ml = st.m / nat
    for c, l, f in zip(ds, ld0, fd0):
        assert np.allclose(pl.get_values(pl.get_local_descriptors(c)), (l - ml) @ st.W, rtol=1e-11, atol=1e-12)
        assert np.allclose(pl.get_values(pl.get_force_descriptors(c)), (f - st.m) @ st.W, rtol=1e-11, atol=1e-12)
        assert pl.get_values(pl.get_local_descriptors(c)).shape == (nat, 3)


# ---------------------------------------------------------------------------------------------------------------
# BASELINE configurations at (or near) full size through size-independent properties
# ---------------------------------------------------------------------------------------------------------------
def test_config1_kdpp_full_size(pl, orc):
    """BASELINE config 1: kDPP(ds, GlobalMean(), DotProduct()), N = 1 000 configurations x 26-dim descriptors
    (examples/DPP-ACE-Si shape). K is compared entry by entry with the restated pair loop; the selector is exercised end to end.
    A degree-2 polynomial kernel of 26-dim features has rank <= 26*27/2 = 351, so the batch size must stay below that."""
    rng = np.random.default_rng(1001)
    n, nat, d = 1000, 32, 26
    mu = 2.0 * rng.standard_normal(d)
    ds = pl.DataSet([pl.Configuration(pl.LocalDescriptors(mu + rng.standard_normal((nat, d)))) for _ in range(n)])
    with pytest.raises(ValueError, match="numerical rank"):
        pl.kDPP(ds[0:400], pl.GlobalMean(), pl.DotProduct(), batch_size=380)
    dpp = pl.kDPP(ds, pl.GlobalMean(), pl.DotProduct(), batch_size=200)
    F = np.stack(pl.compute_features(ds, pl.GlobalMean()))
    ref = orc.symmetric(orc.kernel_matrix(F, orc.DotProduct()))
    assert relerr(dpp.K.L, ref) < RTOL and np.array_equal(dpp.K.L, dpp.K.L.T)
    assert abs(pl.get_inclusion_prob(dpp).sum() - 200.0) < 1e-6
    sub = pl.get_random_subset(dpp, rng=np.random.default_rng(0))
    assert len(sub) == len(set(sub)) == 200 and min(sub) >= 0 and max(sub) < n
    # same eigendecomposition input (K within 1e-11) + same seed -> same indices as a selector built on the oracle's K
    ell_ref = pl.EllEnsemble(ref)
    pl.rescale_(ell_ref, 200)
    assert pl.sample(ell_ref, 200, rng=np.random.default_rng(0)) == sub


def test_config3_full_size_checksums(pl):
    """BASELINE config 3 at full size (R = 2 890 000, K = 500, 11.56 GB): trace(A'WA) = sum_r w_r |a_r|^2, A'W1 column sums and
    A'Wb against independent fp64 torch reductions; exact symmetry; linearity across a 2-way row split."""
    import torch

    from plb200.device import DeviceEngine

    eng = DeviceEngine()
    ncfg, natoms, K = 10000, 96, 500
    ne, nf = ncfg, ncfg * 3 * natoms
    g = torch.Generator(device="cuda").manual_seed(1003)
    A = torch.empty((ne + nf, K), device="cuda", dtype=torch.float64)
    for r0 in range(0, ne + nf, 1 << 19):
        A[r0:r0 + (1 << 19)].normal_(generator=g)
    A[:ne] *= np.sqrt(96.0)
    b = torch.randn(ne + nf, device="cuda", dtype=torch.float64, generator=g)
    we, wf = 100.0, 1.0
    S, vec = eng.syrk_partial(A, None, ne, we, wf, b, None, True)
    torch.cuda.synchronize()
    assert bool(torch.equal(S, S.T))
    w = torch.cat([torch.full((ne,), we, device="cuda", dtype=torch.float64), torch.full((nf,), wf, device="cuda", dtype=torch.float64)])
    tr_ref = float(((A * A).sum(1) * w).sum())
    assert abs(float(torch.trace(S)) - tr_ref) / tr_ref < 1e-12
    v_ref = torch.zeros(K, device="cuda", dtype=torch.float64)
    for r0 in range(0, ne + nf, 1 << 19):
        v_ref += A[r0:r0 + (1 << 19)].T @ (w[r0:r0 + (1 << 19)] * b[r0:r0 + (1 << 19)])
    assert float((vec[:K] - v_ref).abs().max() / v_ref.abs().max()) < 1e-11
    c_ref = we * A[:ne].sum(0)
    assert float((vec[K:2 * K] - c_ref).abs().max() / c_ref.abs().max()) < 1e-12
    assert abs(float(vec[2 * K]) - we * ne) < 1e-6 and abs(float(vec[2 * K + 1]) - float(we * b[:ne].sum())) < 1e-8
    half = ne + nf // 2
    S1, _ = eng.syrk_partial(A[:half], None, ne, we, wf, None, None, False)
    S2, _ = eng.syrk_partial(A[half:], None, 0, we, wf, None, None, False)
    assert float(((S1 + S2) - S).abs().max() / S.abs().max()) < 1e-13
    del A, b, w


def test_config4_slice_checksums(pl):
    """BASELINE config 4 shape (K = 1 000, mean/sigma = 10) on a 1.25 M-row slice (10 GB): mean, trace of the centred covariance
    and a random 64 x 64 sub-block against independent fp64 torch reductions; mean persistence."""
    import torch

    from plb200.device import DeviceEngine
    from plb200.distributed import covariance_sharded

    eng = DeviceEngine()
    n, K = 1_250_000, 1000
    g = torch.Generator(device="cuda").manual_seed(1004)
    shift = 10.0 * torch.randn(K, device="cuda", dtype=torch.float64, generator=g)
    X = torch.empty((n, K), device="cuda", dtype=torch.float64)
    for r0 in range(0, n, 1 << 18):
        X[r0:r0 + (1 << 18)].normal_(generator=g)
        X[r0:r0 + (1 << 18)] += shift
    C, m = covariance_sharded(X, n, engine=eng)
    torch.cuda.synchronize()
    m_ref = X.sum(0) / n
    assert float(((m - m_ref).abs() / m_ref.abs()).max()) < 1e-12
    assert bool(torch.equal(C, C.T))
    idx = torch.randperm(K, device="cuda", generator=g)[:64]
    Xs = X[:, idx] - m_ref[idx]
    blk_ref = (Xs.T @ Xs) / n
    blk = C[idx][:, idx]
    assert float((blk - blk_ref).abs().max() / blk_ref.abs().max()) < 1e-11
    tr_ref = 0.0
    for r0 in range(0, n, 1 << 18):
        Y = X[r0:r0 + (1 << 18)] - m_ref
        tr_ref += float((Y * Y).sum())
    assert abs(float(torch.trace(C)) - tr_ref / n) / (tr_ref / n) < 1e-12
    C2, m2 = covariance_sharded(X, n, mean=m, engine=eng)  # pca.m persists (pca_state.jl:34)
    # (the persisted-mean path centres directly, the first fit went through the one-pass shifted form: same sums, different rounding)
    assert bool(torch.equal(m2, m)) and float((C2 - C).abs().max() / C.abs().max()) < 1e-13 and bool(torch.equal(C2, C2.T))
    del X


def test_config4_full_size_checksums(pl):
    """BASELINE config 4 at FULL size (10^7 rows x K = 1000, 80 GB resident, mean/sigma = 10) through pl_cov_sharded_dev (one rank: no
    communicator, same code path as the 8-GPU run): mean, trace of the centred covariance and a random 64 x 64 sub-block against
    independent chunked fp64 torch reductions; exact symmetry. Skipped when the GPU cannot hold 80 GB + scratch."""
    import torch

    from plb200 import _lib

    torch.cuda.empty_cache()
    _lib.check(_lib.lib().pl_workspace_trim(1), "trim")
    n, K = 10_000_000, 1000
    if torch.cuda.mem_get_info()[0] < (n * K * 8 + 12e9):
        pytest.skip("needs 92 GB of free HBM")
    g = torch.Generator(device="cuda").manual_seed(1004)
    shift = 10.0 * torch.randn(K, device="cuda", dtype=torch.float64, generator=g)
    X = torch.empty((n, K), device="cuda", dtype=torch.float64)
    step = 1 << 18
    for r0 in range(0, n, step):
        X[r0:r0 + step].normal_(generator=g)
        X[r0:r0 + step] += shift
    C = torch.empty((K, K), device="cuda", dtype=torch.float64)
    m = torch.empty(K, device="cuda", dtype=torch.float64)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().pl_cov_sharded_dev(X.data_ptr(), n, n, K, K, m.data_ptr(), 0, 1, C.data_ptr(), st), "pl_cov_sharded_dev")
    torch.cuda.synchronize()
    # independent reference: chunked sums in fp64 (pairwise inside torch, sequential over 39 chunks)
    s = torch.zeros(K, device="cuda", dtype=torch.float64)
    for r0 in range(0, n, step):
        s += X[r0:r0 + step].sum(0)
    m_ref = s / n
    assert float(((m - m_ref).abs() / m_ref.abs()).max()) < 1e-12
    assert bool(torch.equal(C, C.T))
    idx = torch.randperm(K, device="cuda", generator=g)[:64]
    blk_ref = torch.zeros((64, 64), device="cuda", dtype=torch.float64)
    tr_ref = 0.0
    for r0 in range(0, n, step):
        Y = X[r0:r0 + step] - m_ref
        tr_ref += float((Y * Y).sum())
        Ys = Y[:, idx]
        blk_ref += Ys.T @ Ys
        del Y, Ys
    blk_ref /= n
    blk = C[idx][:, idx]
    assert float((blk - blk_ref).abs().max() / blk_ref.abs().max()) < 1e-11
    assert abs(float(torch.trace(C)) - tr_ref / n) / (tr_ref / n) < 1e-12
    del X
    torch.cuda.empty_cache()


def test_config5_shard_sampled_tiles(pl, orc):
    """BASELINE config 5 (DotProduct Gram, N = 200 000, d = 500): shard 0 of 8 (12 214 packed 128 x 128 tiles, 20 GB) exactly as one of the 8
    GPUs computes it. 64 sampled tiles against the ORACLE's C pair loop (kernels.jl:30-36 restated), every diagonal element == 1, and
    2 000 more tiles against an independent torch fp64 evaluation."""
    import ctypes

    import torch

    from plb200 import _lib

    lib = _lib.lib()
    torch.cuda.empty_cache()
    _lib.check(lib.pl_workspace_trim(1), "trim")
    n, d, nparts, part = 200_000, 500, 8, 0
    ntiles = int(lib.pl_gram_tile_count(n, 0, part, nparts))
    if torch.cuda.mem_get_info()[0] < ntiles * 131072 + 4e9:
        pytest.skip("needs 24 GB of free HBM")
    rng = np.random.default_rng(1005)
    mu = torch.as_tensor(2.0 * rng.standard_normal(d), device="cuda")
    g = torch.Generator(device="cuda").manual_seed(1005)
    X = mu + torch.randn((n, d), device="cuda", dtype=torch.float64, generator=g) / np.sqrt(96.0)
    tiles = torch.empty((ntiles, 128, 128), device="cuda", dtype=torch.float64)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.pl_gram_dev(_lib.PL_KERNEL_DOT, X.data_ptr(), n, d, None, 0, 0, d, 2, 0, None, 1.0, 1.0, tiles.data_ptr(), 128, _lib.PL_FILL_FULL, part, nparts, 1,
                               st), "pl_gram_dev")
    torch.cuda.synchronize()
    Xh = X.cpu().numpy()
    T = (n + 127) // 128

    def coords(k):
        ti, tj = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(lib.pl_gram_tile_coords(n, 0, part, nparts, int(k), ctypes.byref(ti), ctypes.byref(tj)), "coords")
        return ti.value, tj.value

    ks = rng.choice(ntiles, 64, replace=False).tolist() + [0, ntiles - 1]
    checked = 0
    for k in ks:
        ti, tj = coords(k)
        if ti < 0:
            continue  # an unused slot of the folded enumeration (odd T)
        r0, r1, c0, c1 = ti * 128, min(ti * 128 + 128, n), tj * 128, min(tj * 128 + 128, n)
        ref = orc.kernel_matrix_cross(Xh[r0:r1], Xh[c0:c1], orc.DotProduct())
        got = tiles[k, : r1 - r0, : c1 - c0].cpu().numpy()
        if ti == tj:  # both triangles of a packed diagonal tile are written; diagonal forced to 1 (n/sqrt(n*n))
            assert np.array_equal(np.diag(got), np.ones(r1 - r0))
            np.fill_diagonal(ref, 1.0)
        assert relerr(got, ref) < RTOL, (k, ti, tj)
        checked += 1
    assert checked >= 60
    # 2 000 further tiles against torch fp64 (normalised rows, DGEMM, square)
    Xn = X / X.norm(dim=1, keepdim=True)
    worst = 0.0
    for k in rng.choice(ntiles, 2000, replace=False).tolist():
        ti, tj = coords(k)
        if ti < 0 or ti == tj or ti == T - 1 or tj == T - 1:
            continue
        ref = (Xn[ti * 128:(ti + 1) * 128] @ Xn[tj * 128:(tj + 1) * 128].T) ** 2
        worst = max(worst, float(((tiles[k] - ref).abs() / ref).max()))
    assert worst < 1e-11, worst
    del tiles, X, Xn
    torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------------------------------
# feature / row packing on the device
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d", [1, 7, 8, 26, 63, 64, 65, 300])
def test_global_features_bit_exact(pl, orc, d):
    """GlobalMean / GlobalSum on the GPU walk Julia's summation tree: bit-identical to the restated reference sum."""
    rng = np.random.default_rng(d)
    natoms = [1, 2, 3, 20, 96, 1023, 1024, 1025, 2500] if d <= 26 else [1, 20, 96, 1030]
    blocks = [rng.standard_normal((n, d)) * 3.0 + rng.standard_normal(d) for n in natoms for _ in range(2)]
    for mean in (True, False):
        X = pl.global_features(blocks, mean=mean)
        Xo = orc.global_features(blocks, mean)
        assert X.shape == (len(blocks), d)
        assert np.array_equal(X, Xo), (d, mean, float(np.max(np.abs(X - Xo))))
    with pytest.raises(ValueError):
        pl.global_features([np.ones((2, d)), np.ones((0, d))])


def test_compute_features_and_kernel_matrix_from_dataset(pl, orc):
    """KernelMatrix(ds, GlobalMean(), k) (kernels.jl:250-253): features on the GPU, then the Gram kernel."""
    rng = np.random.default_rng(12)
    ds = fake_dataset(pl, rng, 40, 20, 8)
    F = pl.compute_features(ds, pl.GlobalMean())
    Fo = [orc.global_mean(list(pl.get_values(pl.get_local_descriptors(c)))) for c in ds]
    assert np.array_equal(F, np.stack(Fo))
    S = pl.compute_features(ds, pl.GlobalSum())
    assert np.array_equal(S, np.stack([orc.global_sum(list(pl.get_values(pl.get_local_descriptors(c)))) for c in ds]))
    K = np.asarray(pl.KernelMatrix(ds, pl.GlobalMean(), pl.RBF(pl.Euclidean(8))))
    Ko = orc.symmetric(orc.kernel_matrix(Fo, orc.RBF(orc.Euclidean(8))))
    assert relerr(K, Ko) < RTOL


def test_global_features_device_resident_bandwidth_shape(pl, orc):
    """device-resident form on torch tensors (config-2 shape scaled down: 2000 configurations x 96 atoms x 300)."""
    import torch

    rng = np.random.default_rng(5)
    N, nat, d = 2000, 96, 300
    L = rng.standard_normal((N * nat, d))
    off = np.arange(N + 1, dtype=np.int64) * nat
    Ld, offd = torch.from_numpy(L).cuda(), torch.from_numpy(off).cuda()
    X = pl.global_features_dev(Ld, offd, mean=True)
    torch.cuda.synchronize()
    Xo = orc.global_features([L[i * nat:(i + 1) * nat] for i in range(0, N, 97)], True)
    assert np.array_equal(X.cpu().numpy()[::97], Xo)


@pytest.mark.parametrize("chunk_rows", [0, 16, 64, 1000])
def test_normal_equations_blocks_streamed(pl, orc, chunk_rows):
    """pl_normal_eq_blocks: lp.B / lp.dB[i] handed over as they lie in memory, rows streamed in chunks (any chunking must
    give the same answer to rounding; the chunk partials are accumulated in fixed order)."""
    rng = np.random.default_rng(33)
    K, ncfg = 24, 37
    natoms = rng.integers(1, 9, size=ncfg)
    B = [rng.standard_normal(K) * 4 for _ in range(ncfg)]
    dB = [np.asfortranarray(rng.standard_normal((K, 3 * n))) for n in natoms]  # Julia's column-major K x 3natoms
    e = rng.standard_normal(ncfg)
    f = [rng.standard_normal(3 * n) for n in natoms]
    A = np.concatenate([np.stack(B)] + [m.T for m in dB], axis=0)
    b = np.concatenate([e] + f)
    R = A.shape[0]
    for intercept, lam, (we, wf) in ((False, 0.0, (1.0, 1.0)), (True, 0.01, (100.0, 1.0))):
        M, v = pl.normal_equations_blocks(B + dB, b, None, ncfg, we, wf, intercept, lam, chunk_rows=chunk_rows)
        q = np.concatenate([we * np.ones(ncfg), wf * np.ones(R - ncfg)])
        Ai = np.hstack([np.concatenate([np.ones(ncfg), np.zeros(R - ncfg)])[:, None], A]) if intercept else A
        Mo, vo = orc.normal_equations(Ai, b, q, lam)
        assert normerr(M, Mo) < RTOL and normerr(v, vo) < RTOL
        assert np.array_equal(M, M.T)
    w = rng.uniform(0.1, 3.0, size=R)
    M, v = pl.normal_equations_blocks(B + dB, b, w, chunk_rows=chunk_rows)
    Mo, vo = orc.normal_equations(A, b, w, 0.0)
    assert normerr(M, Mo) < RTOL and normerr(v, vo) < RTOL
    # the same call twice is bitwise reproducible
    M2, v2 = pl.normal_equations_blocks(B + dB, b, w, chunk_rows=chunk_rows)
    assert np.array_equal(M, M2) and np.array_equal(v, v2)


def test_normal_equations_large_input_streams(pl, orc):
    """pl_normal_eq switches to the streamed path above 768 MB of rows; same result as the resident device path."""
    import torch

    rng = np.random.default_rng(8)
    R, K = 420_000, 256  # 860 MB
    A = rng.standard_normal((R, K))
    b = rng.standard_normal(R)
    M, v = pl.normal_equations(A, b, None, 1000, 30.0, 1.0, True, 0.01)
    from plb200.device import DeviceEngine

    eng = DeviceEngine()
    S, vec = eng.syrk_partial(torch.from_numpy(A).cuda(), None, 1000, 30.0, 1.0, torch.from_numpy(b).cuda(), None, True)
    Md, vd = eng.normal_eq_finish(S, vec, True, 0.01)
    assert normerr(M, Md.cpu().numpy()) < 1e-12 and normerr(v, vd.cpu().numpy()) < 1e-12  # chunked vs one-pass summation order
    assert np.array_equal(M, M.T)


# ---------------------------------------------------------------------------------------------------------------
# eigendecomposition and the L-ensemble on the device
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 7, 130, 1000])
def test_eigh_matches_lapack(pl, orc, n):
    """pl_eigh against LAPACK (numpy.linalg.eigh — the same syev family the reference reaches through eigen(Symmetric))."""
    from plb200.dpp import eigh

    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, n + 3))
    A = B @ B.T / (n + 3) + np.diag(rng.uniform(0, 2, n))
    lam, U = eigh(A)
    lo, _ = np.linalg.eigh(A)
    scale = np.max(np.abs(lo))
    assert np.max(np.abs(lam - lo)) / scale < 1e-13
    assert np.all(np.diff(lam) >= 0), "ascending, as eigen() returns"
    assert np.max(np.abs(A @ U - U * lam[None, :])) / scale < 1e-12, "residual"
    assert np.max(np.abs(U.T @ U - np.eye(n))) < 1e-12, "orthonormal"
    # one triangle is enough: Symmetric(:U) storage (the other triangle holds garbage)
    Au = np.triu(A) + np.tril(rng.standard_normal((n, n)), -1)
    lam_u = eigh(Au, fill="U", vectors=False)
    assert np.max(np.abs(lam_u - lo)) / scale < 1e-13


def test_ell_ensemble_resident_pipeline(pl, orc):
    """EllEnsemble(KernelMatrix(F, k)) fused on the device: λ, K, inclusion probabilities and selected eigenvectors against the
    host-side definitions evaluated on the oracle's kernel matrix."""
    rng = np.random.default_rng(77)
    n, d = 700, 12
    X = rng.standard_normal((n, d)) * 0.4 + rng.standard_normal(d)
    k = pl.RBF(pl.Euclidean(d), ell=1.5)
    ell = pl.EllEnsemble.from_features(X, k)
    Ko = orc.symmetric(orc.kernel_matrix(X, orc.RBF(orc.Euclidean(d), ell=1.5)))
    assert relerr(ell.L, Ko) < RTOL
    lo, Uo = np.linalg.eigh(Ko)
    assert np.max(np.abs(ell.λ - np.maximum(lo, 0))) / lo.max() < 1e-12
    pl.rescale_(ell, 50)
    p = pl.inclusion_prob(ell)  # device: U never downloaded
    assert ell._U is None
    g = ell.α * np.maximum(lo, 0)
    g = g / (1 + g)
    po = np.sum(Uo * Uo * g[None, :], axis=1)
    assert abs(p.sum() - 50.0) < 1e-8 and np.max(np.abs(p - po)) < 1e-10
    J = [n - 1, n - 2, 5, n - 40]
    V = ell.eigvecs(J)
    assert V.shape == (n, 4) and ell._U is None
    for c, jj in enumerate(J):
        if lo[jj] > 1e-6 * lo.max():  # (eigenvectors of the numerically-null space are not unique)
            assert np.max(np.abs(Ko @ V[:, c] - lo[jj] * V[:, c])) / lo.max() < 1e-11
    s1 = pl.sample(ell, 50, rng=np.random.default_rng(3))
    assert len(set(s1)) == 50
    # a second resident ensemble evicts the first one, which stays usable (small: U is downloaded on eviction)
    ell2 = pl.EllEnsemble.from_features(X[:100], pl.DotProduct())
    assert ell._U is not None and not ell._resident and ell2._resident
    assert pl.sample(ell, 50, rng=np.random.default_rng(3)) == s1


def test_compute_eigen_and_pca_on_device(pl, orc, monkeypatch):
    """compute_eigen (dimension_reduction.jl:11-14): Q on the GPU; eigen(Symmetric(Q)) stays LAPACK by default (same solver as the reference:
    identical vectors for an identical Q), the device eigensolver is opt-in (PLB200_EIGEN=gpu). Eigenvalues against the oracle."""
    rng = np.random.default_rng(4)
    rows = rng.standard_normal((500, 30)) @ np.diag(np.linspace(0.1, 3, 30))
    lo, _ = orc.compute_eigen(list(rows))
    Q = rows.T @ rows / 500
    launches = pl.launch_count()
    lam, phi = pl.compute_eigen(rows)
    Qg, _ = pl.covariance(rows)
    lam_l, phi_l = np.linalg.eigh(Qg, UPLO="U")
    assert np.array_equal(lam, lam_l) and np.array_equal(phi, phi_l), "default: LAPACK on the GPU-built Q"
    monkeypatch.setenv("PLB200_EIGEN", "gpu")
    lam_g, phi_g = pl.compute_eigen(rows)
    for l, p in ((lam, phi), (lam_g, phi_g)):
        assert np.max(np.abs(l - lo)) / lo.max() < 1e-12
        assert np.max(np.abs(Q @ p - p * l[None, :])) / lo.max() < 1e-12
    assert pl.launch_count() > launches


def test_centred_syrk_with_uniform_weight_and_ragged_rows(pl):
    """pl_syrk_dev documents S = sum_r w_r (a_r - m)(a_r - m)'. With a uniform weight != 1 and R % 16 != 0 the ragged last stage must mask the
    zero-filled rows without applying the weight a second time."""
    import torch

    from plb200.device import DeviceEngine

    eng = DeviceEngine()
    g = torch.Generator(device="cuda").manual_seed(5)
    for R, K in ((1000 + 7, 40), (16 * 70 + 1, 130), (33, 8)):
        A = torch.randn((R, K), device="cuda", dtype=torch.float64, generator=g) + 3.0
        m = A.mean(0)
        for (ne, we, wf) in ((0, 1.0, 2.5), (R, 0.3, 1.0), (R // 2, 2.0, 2.0)):
            S, _ = eng.syrk_partial(A, None, ne, we, wf, None, m, False)
            w = torch.full((R,), wf, device="cuda", dtype=torch.float64)
            w[:ne] = we
            Ac = A - m
            ref = Ac.T @ (Ac * w[:, None])
            assert float((S - ref).abs().max() / ref.abs().max()) < 1e-12, (R, K, ne, we, wf)


def test_normal_equations_aug_and_fallback_agree(pl, orc, monkeypatch):
    """The moment vectors through the augmented columns (default) and through the weighted column-sum pass (PLB200_SYRK_NO_AUG)
    are two routes to the same sums: they must agree to rounding, and S itself must be bit-identical (the main contraction is
    untouched by the extra columns)."""
    import os

    rng = np.random.default_rng(91)
    R, K, ne = 30000, 500, 1200
    A = rng.standard_normal((R, K)) + 0.3
    b = rng.standard_normal(R)
    w = rng.uniform(0.1, 3.0, size=R)
    for ww in (None, w):
        M1, v1 = pl.normal_equations(A, b, ww, ne, 100.0, 1.0, True, 0.0)
        os.environ["PLB200_SYRK_NO_AUG"] = "1"
        try:
            M2, v2 = pl.normal_equations(A, b, ww, ne, 100.0, 1.0, True, 0.0)
        finally:
            del os.environ["PLB200_SYRK_NO_AUG"]
        assert np.array_equal(M1[1:, 1:], M2[1:, 1:])
        assert normerr(M1, M2) < 1e-12 and normerr(v1, v2) < 1e-12  # two summation orders of the same 30 000 terms


def test_golden_vectors_through_the_c_abi(pl):
    """The committed extended-precision golden vectors (tests/golden/hotpath_golden.npz) straight against the CUDA path."""
    import os

    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hotpath_golden.npz"))
    F = G["c1_feat"]
    assert relerr(np.asarray(pl.KernelMatrix(F, pl.DotProduct(3))), G["c1_dot3"]) < RTOL
    assert relerr(np.asarray(pl.KernelMatrix(F, pl.RBF(pl.Euclidean(26), ell=0.8, beta=2.0))), G["c1_rbf"]) < RTOL
    assert relerr(np.asarray(pl.KernelMatrix(F, pl.RBF(pl.Euclidean(pl.Diagonal(G["c1_cdiag"])), ell=1.5))), G["c1_rbf_diag"]) < RTOL
    assert relerr(np.asarray(pl.KernelMatrix(F, pl.RBF(pl.Euclidean(pl.Symmetric(G["c1_cfull"])), ell=2.0))), G["c1_rbf_full"]) < RTOL
    assert relerr(np.asarray(pl.KernelMatrix(F, pl.InverseMultiquadric(pl.Euclidean(26), c2=1.3, ell=0.9))), G["c1_imq"]) < RTOL
    A, b, q = G["ne_A"], G["ne_b"], G["ne_q"]
    M, v = pl.normal_equations(A, b, None, 12, 100.0, 1.0, False, 0.0)
    assert normerr(M, G["ne_M"]) < RTOL and normerr(v, G["ne_v"]) < RTOL
    M, v = pl.normal_equations(A, b, q, 12, 1.0, 1.0, True, 0.01)  # the same weights as a per-row array, with intercept and ridge
    assert normerr(M, G["ne_Mi"]) < RTOL and normerr(v, G["ne_vi"]) < RTOL
    C, m = pl.covariance(G["cov_X"], center=True)
    assert relerr(m, G["cov_mean"]) < 1e-13 and normerr(C, G["cov_centered"]) < RTOL
    C0, _ = pl.covariance(G["cov_X"], center=False)
    assert normerr(C0, G["cov_raw"]) < RTOL
    off, L = G["gf_offsets"], G["gf_local"]
    assert relerr(pl.global_features(None, mean=False, stacked=(L, off)), G["gf_sum"]) < 1e-13
    assert relerr(pl.global_features(None, mean=True, stacked=(L, off)), G["gf_mean"]) < 1e-13
    y = pl.predict(A, G["pr_beta"], 0.37, 12)
    assert normerr(y, G["pr_y"]) < RTOL


@pytest.mark.parametrize("n2", [1, 5, 31, 32, 33, 64, 96, 97, 130])
def test_narrow_tile_cross_kernels_and_projections(pl, orc, n2):
    """Rectangular results with few columns run on the 128 x 32 tile (n2 <= 96) instead of the 128 x 128 one: every kernel
    against the oracle across the switch-over, plus the projection W'(x - m) of transform! (pca_state.jl:42-62)."""
    rng = np.random.default_rng(1000 + n2)
    m, d = 300, 37
    mu = 2.0 * rng.standard_normal(d)  # a common mean: cosines stay away from 0 (a relative test of cos^3 near 0 measures cancellation)
    F1 = rng.standard_normal((m, d)) * 0.5 + mu
    F2 = rng.standard_normal((n2, d)) * 0.5 + mu
    cd = rng.uniform(0.5, 2.0, size=d)
    for kp, ko in ((pl.DotProduct(3), orc.DotProduct(3)),
                   (pl.RBF(pl.Euclidean(d), ell=1.7, beta=0.5), orc.RBF(orc.Euclidean(d), ell=1.7, beta=0.5)),
                   (pl.RBF(pl.Euclidean(pl.Diagonal(cd)), ell=2.0), orc.RBF(orc.Euclidean(d, cd), ell=2.0)),
                   (pl.InverseMultiquadric(pl.Euclidean(d), c2=0.7, ell=1.1), orc.InverseMultiquadric(orc.Euclidean(d), c2=0.7, ell=1.1))):
        K = pl.KernelMatrix(F1, F2, kp)
        Ko = orc.kernel_matrix_cross(F1, F2, ko, use_c=True)
        assert K.shape == (m, n2) and relerr(K, Ko) < RTOL, type(kp).__name__
    # projection: X (rows x K) times W (K x p), shifted
    X = rng.standard_normal((1000, 64)) + 3.0
    W = np.linalg.qr(rng.standard_normal((64, min(n2, 64))))[0]
    shift = X.mean(0)
    Y = pl.project(X, W, shift)
    assert normerr(Y, (X - shift) @ W) < 1e-11  # the shift is applied as X·W - shift·W (|shift|/sigma = 3)


def test_no_out_of_bounds_writes(pl):
    """Guard-band check of every device entry point that writes a caller buffer (compute-sanitizer is not available on the
    pool): outputs are views into larger sentinel-filled allocations; after the call the sentinels must be intact —
    leading/trailing guard regions and, for matrices, the padding columns between n and the leading dimension."""
    import torch

    from plb200 import _lib

    lib = _lib.lib()
    dev = torch.device("cuda")
    SENT = -777.25
    st = torch.cuda.current_stream().cuda_stream

    def guarded(rows, cols, ld):
        off = (ld + 17) & ~1  # even element offset: the view stays 16-byte aligned
        buf = torch.full((off + (rows + 1) * ld + 64,), SENT, device=dev, dtype=torch.float64)
        view = buf[off: off + rows * ld].view(rows, ld)
        return buf, view

    def check(buf, view, rows, cols, ld, what, written_mask=None):
        torch.cuda.synchronize()
        off = (ld + 17) & ~1
        b = buf.clone()
        v = b[off: off + rows * ld].view(rows, ld)
        inner = v[:, :cols]
        assert torch.all(v[:, cols:] == SENT), f"{what}: padding columns overwritten"
        if written_mask is not None:
            assert torch.all(inner[~written_mask] == SENT), f"{what}: wrote outside the requested triangle"
        inner.fill_(SENT)
        assert torch.all(b == SENT), f"{what}: wrote outside the output matrix"

    g = torch.Generator(device=dev).manual_seed(11)
    for n, d in ((1, 3), (129, 26), (300, 300), (517, 77)):
        X = torch.randn((n, d + (d & 1)), device=dev, dtype=torch.float64, generator=g)[:, :d]
        for kid in (_lib.PL_KERNEL_DOT, _lib.PL_KERNEL_RBF, _lib.PL_KERNEL_IMQ):
            for fill in (_lib.PL_FILL_FULL, _lib.PL_FILL_RM_UPPER, _lib.PL_FILL_RM_LOWER):
                ld = (n + 7) & ~1
                buf, out = guarded(n, n, ld)
                _lib.check(lib.pl_gram_dev(kid, X.data_ptr(), n, X.stride(0), None, 0, 0, d, 2, 0, None, 1.0, 1.0, out.data_ptr(), ld, fill, 0, 1, 0, st), "pl_gram_dev")
                # inside diagonal 128 x 128 tiles the unrequested half is written as 0 (zeros(n,n) storage); whole tiles of the
                # unrequested triangle must stay untouched
                ii, jj = torch.meshgrid(torch.arange(n, device=dev), torch.arange(n, device=dev), indexing="ij")
                same_tile = (ii // 128) == (jj // 128)
                mask = torch.ones((n, n), dtype=torch.bool, device=dev) if fill == _lib.PL_FILL_FULL else (
                    ((jj >= ii) | same_tile) if fill == _lib.PL_FILL_RM_UPPER else ((jj <= ii) | same_tile))
                check(buf, out, n, n, ld, f"gram sym kid={kid} fill={fill} n={n}", mask)
        for n2 in (1, 33, 97, 200):  # narrow and wide rectangular tiles
            X2 = torch.randn((n2, d + (d & 1)), device=dev, dtype=torch.float64, generator=g)[:, :d]
            ld = (n2 + 5) & ~1
            buf, out = guarded(n, n2, ld)
            _lib.check(lib.pl_gram_dev(_lib.PL_KERNEL_RBF, X.data_ptr(), n, X.stride(0), X2.data_ptr(), n2, X2.stride(0), d, 0, 0, None, 1.3, 1.0,
                                       out.data_ptr(), ld, 3, 0, 1, 0, st), "pl_gram_dev")
            check(buf, out, n, n2, ld, f"gram cross n={n} n2={n2}")
    for R, K, ne in ((1, 1, 1), (1000, 127, 10), (5000, 500, 100), (2049, 128, 0), (700, 1000, 700)):
        A = torch.randn((R, K + (K & 1)), device=dev, dtype=torch.float64, generator=g)[:, :K]
        b = torch.randn(R, device=dev, dtype=torch.float64, generator=g)
        m = A.mean(0).contiguous()
        for centred in (False, True):
            bufS, S = guarded(K, K, K)
            bufv, vec = guarded(1, 2 * K + 2, 2 * K + 2)
            _lib.check(lib.pl_syrk_dev(A.data_ptr(), R, K, A.stride(0), None, ne, 30.0, 1.0, None if centred else b.data_ptr(),
                                       m.data_ptr() if centred else None, S.data_ptr(), None if centred else vec.data_ptr(), st), "pl_syrk_dev")
            check(bufS, S, K, K, K, f"syrk S R={R} K={K} centred={centred}")
            check(bufv, vec, 1, 2 * K + 2, 2 * K + 2, f"syrk vec R={R} K={K}")
        bufy, y = guarded(1, R, R)
        beta = torch.randn(K, device=dev, dtype=torch.float64, generator=g)
        _lib.check(lib.pl_predict_dev(A.data_ptr(), R, K, A.stride(0), beta.data_ptr(), 0.5, ne, y.data_ptr(), st), "pl_predict_dev")
        check(bufy, y, 1, R, R, f"predict R={R} K={K}")
        bufc, cs = guarded(1, K, K)
        _lib.check(lib.pl_colsum_dev(A.data_ptr(), R, K, A.stride(0), cs.data_ptr(), st), "pl_colsum_dev")
        check(bufc, cs, 1, K, K, f"colsum R={R} K={K}")
    for N, nat, d in ((1, 1, 1), (37, 5, 26), (10, 1100, 7), (300, 3, 301)):
        L = torch.randn((N * nat, d + (d & 1)), device=dev, dtype=torch.float64, generator=g)[:, :d]
        off = (torch.arange(N + 1, device=dev, dtype=torch.int64) * nat).contiguous()
        ld = (d + 4) & ~1
        buf, Xo = guarded(N, d, ld)
        _lib.check(lib.pl_global_feature_dev(L.data_ptr(), L.stride(0), off.data_ptr(), N, d, 1, Xo.data_ptr(), ld, N * nat, st), "pl_global_feature_dev")
        check(buf, Xo, N, d, ld, f"global_feature N={N} d={d}")


def test_out_of_core_accumulator(pl, orc):
    """pl_neq_begin / pl_neq_add / pl_neq_finish: ooc_learn!'s running sums (linear-learn.jl:355-356). Any batching gives the one-shot
    result to rounding, a fixed batching is bitwise reproducible, finish is non-destructive, and a generator of configurations works."""
    rng = np.random.default_rng(17)
    K = 40
    natoms = rng.integers(2, 30, size=120)

    def gen():
        r = np.random.default_rng(18)
        for n in natoms:
            yield (r.standard_normal(K) * n, r.standard_normal((3 * n, K)), r.standard_normal() * n, r.standard_normal(3 * n))

    configs = list(gen())
    Ao, bo = orc.ooc_accumulate(configs, ws=(30.0, 1.0), eweight_normalized="squared")
    results = []
    for batch_rows in (1, 100, 1 << 16):
        beta = np.zeros(K)
        M, v = pl.ooc_learn_(beta, gen(), ws=(30.0, 1.0), λ=None, eweight_normalized="squared", batch_rows=batch_rows)
        assert normerr(M, Ao) < RTOL and normerr(v, bo) < RTOL
        results.append((M, v))
    M2, v2 = pl.ooc_learn_(np.zeros(K), gen(), ws=(30.0, 1.0), λ=None, eweight_normalized="squared", batch_rows=100)
    assert np.array_equal(M2, results[1][0]) and np.array_equal(v2, results[1][1])
    acc = pl.NormalEquationAccumulator(K, batch_rows=64)
    for gd, fdm, e, f in configs[:50]:
        acc.add(np.concatenate([gd[None, :], fdm]), np.concatenate([[e], f]), np.ones(1 + fdm.shape[0]))
    Ma, va = acc.result()
    Mb, vb = acc.result(intercept=False, lam=0.5)  # finish twice: the accumulators stay open
    assert np.array_equal(Mb - 0.5 * np.eye(K), Ma) or normerr(Mb - 0.5 * np.eye(K), Ma) < 1e-15
    for gd, fdm, e, f in configs[50:]:
        acc.add(np.concatenate([gd[None, :], fdm]), np.concatenate([[e], f]), np.ones(1 + fdm.shape[0]))
    Mc, vc = acc.result()
    A = np.concatenate([np.concatenate([gd[None, :], fdm]) for gd, fdm, e, f in configs])
    b = np.concatenate([np.concatenate([[e], f]) for gd, fdm, e, f in configs])
    assert normerr(Mc, A.T @ A) < RTOL and normerr(vc, A.T @ b) < RTOL


def test_correlation_features_bit_exact_ragged(pl, orc):
    """pl_correlation_feature over ragged configurations (1 .. 300 atoms): bit-identical to the reference-order evaluation."""
    rng = np.random.default_rng(23)
    for d in (1, 5, 26):
        blocks = [rng.standard_normal((n, d)) * 2.0 + rng.standard_normal(d) for n in (1, 2, 3, 32, 300)]
        C = pl.correlation_features(blocks)
        for c, b in zip(C, blocks):
            ref = orc.correlation_matrix(list(b))
            assert np.array_equal(np.asarray(c), ref) and np.array_equal(c.data, c.data.T)


@pytest.mark.parametrize("n,d", [(2, 1), (30, 8), (150, 26), (300, 7)])
def test_kernel_stein_discrepancy(pl, orc, n, d):
    """compute_divergence(x, ::KernelSteinDiscrepancy) (divergences.jl:27-49) through pl_ksd_rbf against the literal pair loop,
    for identity / Diagonal / Symmetric Cinv, plus the reference's own test (kernel_tests.jl:93-97)."""
    rng = np.random.default_rng(n * 100 + d)
    x = [rng.standard_normal(d) * 0.8 + 5.0 for _ in range(n)]  # mean-heavy on purpose: the Gram forms must not cancel
    score = lambda v: -(v - 5.0)
    cd = rng.uniform(0.5, 2.0, size=d)
    B = rng.standard_normal((d, d))
    cf = B @ B.T / d + np.eye(d)
    for ep, eo in ((pl.Euclidean(d), orc.Euclidean(d)), (pl.Euclidean(pl.Diagonal(cd)), orc.Euclidean(d, cd)),
                   (pl.Euclidean(pl.Symmetric(cf)), orc.Euclidean(d, cf))):
        for ell, beta in ((1.0, 1.0), (2.5, 0.7)):
            got = pl.compute_divergence(x, pl.KernelSteinDiscrepancy(score=score, kernel=pl.RBF(ep, ell=ell, beta=beta)))
            ref = orc.compute_divergence_ksd(x, score, orc.RBF(eo, ell=ell, beta=beta))
            assert abs(got - ref) <= 1e-10 * max(abs(ref), 1e-3), (got, ref)
    if n >= 30:
        f_mvn = [rng.standard_normal(d) for _ in range(n)]
        f_gm = [rng.standard_normal(d) + (3.0 if i % 2 else -3.0) for i in range(n)]
        ksd = pl.KernelSteinDiscrepancy(score=lambda v: -v, kernel=pl.RBF(pl.Euclidean(d)))
        assert isinstance(ksd, pl.Divergence)
        assert pl.compute_divergence(f_mvn, ksd) < pl.compute_divergence(f_gm, ksd)
    with pytest.raises(TypeError):
        pl.compute_divergence(x, pl.KernelSteinDiscrepancy(score=score, kernel=pl.DotProduct()))


def test_minibatch_assembly_sums(pl, orc):
    """learn!(lp, ss::SubsetSelector, α) (linear-learn.jl:92-95, 137-143): the per-step sums over a random subset of configurations."""
    rng = np.random.default_rng(41)
    ds = fake_dataset(pl, rng, 60, 7, 8, energies=True, forces=True)
    lp = pl.LinearProblem(ds)
    inds = rng.choice(60, size=25, replace=True)  # get_random_subset may repeat indices (random.jl / dbscan.jl sample with replacement)
    AtAe, Atbe, AtAf, Atbf = pl.batch_sums(lp, inds)
    re, rbe = orc.ols_sums([lp.B[i] for i in inds], [lp.e[i] for i in inds])
    rf, rbf = orc.ols_sums([lp.dB[i] for i in inds], [lp.f[i] for i in inds])
    assert normerr(AtAe, re) < RTOL and normerr(Atbe, rbe) < RTOL and normerr(AtAf, rf) < RTOL and normerr(Atbf, rbf) < RTOL
    ds_e = fake_dataset(pl, rng, 30, 5, 6, energies=True)
    lpu = pl.LinearProblem(ds_e)
    AtA, Atb = pl.batch_sums(lpu, [3, 3, 7, 29])
    ra, rb = orc.ols_sums([lpu.iv_data[i] for i in (3, 3, 7, 29)], [lpu.dv_data[i] for i in (3, 3, 7, 29)])
    assert normerr(AtA, ra) < RTOL and normerr(Atb, rb) < RTOL


def test_randomized_shape_sweep(pl):
    """Seeded sweep over odd shapes (every tile / box / split boundary gets hit sooner or later): all kernel families against
    vectorised fp64 numpy evaluations of the reference formulas. Tolerance 1e-10 (the contract); typical errors are 1e-14."""
    rng = np.random.default_rng(20261020)

    def rbf_ref(X1, X2, C, ell, beta):
        U = X1[:, None, :] - X2[None, :, :]
        d2 = np.einsum("ijk,kl,ijl->ij", U, C, U)
        return beta * np.exp(-0.5 * d2 / ell ** 2)

    for trial in range(40):
        n = int(rng.choice([1, 2, 7, 127, 128, 129, 255, 257, 300, 511]))
        d = int(rng.choice([1, 2, 3, 15, 16, 17, 31, 33, 64, 65]))
        mu = 2.0 * rng.standard_normal(d)
        X = rng.standard_normal((n, d)) * 0.4 + mu
        ell, beta = float(rng.uniform(0.5, 3.0)), float(rng.uniform(0.2, 2.0))
        kind = trial % 3
        if kind == 0:
            e, C = pl.Euclidean(d), np.eye(d)
        elif kind == 1:
            cd = rng.uniform(0.5, 2.0, size=d)
            e, C = pl.Euclidean(pl.Diagonal(cd)), np.diag(cd)
        else:
            B = rng.standard_normal((d, d))
            C = B @ B.T / d + np.eye(d)
            e = pl.Euclidean(pl.Symmetric(C))
        K = np.asarray(pl.KernelMatrix(X, pl.RBF(e, ell=ell, beta=beta)))
        assert relerr(K, rbf_ref(X, X, C, ell, beta)) < 1e-10, ("rbf", n, d, kind)
        assert np.array_equal(K, K.T)
        G = X @ X.T
        nr = np.sqrt(np.diag(G))
        alpha = int(rng.choice([1, 2, 3]))
        Kd = np.asarray(pl.KernelMatrix(X, pl.DotProduct(alpha)))
        assert np.max(np.abs(Kd - (G / np.outer(nr, nr)) ** alpha)) < 1e-10, ("dot", n, d, alpha)
        n2 = int(rng.choice([1, 3, 32, 33, 96, 97, 200]))
        X2 = rng.standard_normal((n2, d)) * 0.4 + mu
        Kx = pl.KernelMatrix(X, X2, pl.RBF(e, ell=ell, beta=beta))
        assert relerr(Kx, rbf_ref(X, X2, C, ell, beta)) < 1e-10, ("rbf cross", n, n2, d, kind)
        Ki = np.asarray(pl.KernelMatrix(X, pl.InverseMultiquadric(e, c2=beta, ell=ell)))
        U = X[:, None, :] - X[None, :, :]
        assert relerr(Ki, (beta + np.einsum("ijk,kl,ijl->ij", U, C, U) / ell ** 2) ** -0.5) < 1e-10, ("imq", n, d, kind)
    for trial in range(40):
        R = int(rng.choice([1, 5, 15, 16, 17, 255, 1023, 1024, 1025, 4097]))
        Kf = int(rng.choice([1, 2, 14, 15, 16, 17, 63, 127, 128, 129, 143, 255, 256, 300]))
        ne = int(rng.integers(0, R + 1))
        A = rng.standard_normal((R, Kf)) + 0.5
        b = rng.standard_normal(R)
        intercept = bool(trial % 2)
        lam = float(rng.choice([0.0, 0.01]))
        we, wf = float(rng.uniform(0.5, 100.0)), float(rng.uniform(0.5, 2.0))
        w = rng.uniform(0.1, 3.0, size=R) if trial % 3 == 0 else None
        M, v = pl.normal_equations(A, b, w, ne, we, wf, intercept, lam)
        q = w if w is not None else np.concatenate([we * np.ones(ne), wf * np.ones(R - ne)])
        Ai = np.hstack([np.concatenate([np.ones(ne), np.zeros(R - ne)])[:, None], A]) if intercept else A
        Mo = Ai.T @ (q[:, None] * Ai) + lam * np.eye(Ai.shape[1])
        vo = Ai.T @ (q * b)
        assert np.max(np.abs(M - Mo)) / np.max(np.abs(Mo)) < 1e-10, ("neq", R, Kf, ne, intercept)
        assert np.max(np.abs(v - vo)) / max(np.max(np.abs(vo)), 1e-300) < 1e-10, ("neq rhs", R, Kf, ne)
        assert np.array_equal(M, M.T)
        Xc = A + 10.0 * rng.standard_normal(Kf)
        Cc, m = pl.covariance(Xc, center=True)
        mo = Xc.mean(0)
        Co = (Xc - mo).T @ (Xc - mo) / R
        assert np.max(np.abs(m - mo) / np.abs(mo)) < 1e-12
        if R > 1:
            assert np.max(np.abs(Cc - Co)) / np.max(np.abs(Co)) < 1e-10, ("cov", R, Kf)
        assert np.array_equal(Cc, Cc.T)


def test_degenerate_inputs_propagate_like_the_reference(pl, orc):
    """A zero feature vector makes DotProduct 0/sqrt(0) = NaN in the reference (kernels.jl:35), its own diagonal entry included; NaN
    inputs propagate through every kernel. The GPU path must not paper over either (the diagonal is only forced to its exact value
    where the reference value is finite)."""
    rng = np.random.default_rng(3)
    X = rng.standard_normal((200, 9)) + 1.0
    X[17] = 0.0
    with np.errstate(all="ignore"):
        Ko = orc.symmetric(orc.kernel_matrix(X, orc.DotProduct(), use_c=False))
    K = np.asarray(pl.KernelMatrix(X, pl.DotProduct()))
    assert np.isnan(Ko[17, 17]) and np.isnan(K[17]).all() and np.isnan(K[:, 17]).all()
    ok = ~np.isnan(Ko)
    assert np.array_equal(np.isnan(K), np.isnan(Ko)) and relerr(K[ok], Ko[ok]) < RTOL
    Kr = np.asarray(pl.KernelMatrix(X, pl.RBF(pl.Euclidean(9))))  # a zero vector is an ordinary point for a distance kernel
    assert relerr(Kr, orc.symmetric(orc.kernel_matrix(X, orc.RBF(orc.Euclidean(9)), use_c=False))) < RTOL and Kr[17, 17] == 1.0
    Xn = X.copy()
    Xn[17] = 1.0
    Xn[5, 3] = np.nan
    Kn = np.asarray(pl.KernelMatrix(Xn, pl.DotProduct()))
    assert np.isnan(Kn[5]).all() and np.isnan(Kn[:, 5]).all() and not np.isnan(np.delete(np.delete(Kn, 5, 0), 5, 1)).any()


def test_pinned_result_buffer_reuse(pl, orc):
    """KernelMatrix(F, k, out=pinned_zeros(...)): a caller-owned pinned result buffer reused across calls gives the same matrix."""
    rng = np.random.default_rng(9)
    X = rng.standard_normal((300, 12)) * 0.5 + 1.0
    buf = pl.pinned_zeros((300, 300))
    assert buf.shape == (300, 300) and not buf.any()
    K1 = pl.KernelMatrix(X, pl.RBF(pl.Euclidean(12)), out=buf)
    ref = np.asarray(pl.KernelMatrix(X, pl.RBF(pl.Euclidean(12))))
    assert np.array_equal(np.asarray(K1), ref)
    K2 = pl.KernelMatrix(X[::-1].copy(), pl.DotProduct(), out=buf)  # reuse: the buffer is re-zeroed, no stale triangle survives
    assert np.array_equal(np.asarray(K2), np.asarray(pl.KernelMatrix(X[::-1].copy(), pl.DotProduct())))
    with pytest.raises(ValueError):
        pl.KernelMatrix(X, pl.DotProduct(), out=np.zeros((5, 5)))
    del K1, K2, buf


def test_example_workflow_runs_end_to_end(pl):
    """examples/dpp_subsample_fit_pca.py: the reference examples' workflow (kDPP subsample -> learn! -> predictions -> PCAState
    fit!/transform!) on synthetic descriptors; the linear model must be recovered from the DPP subset."""
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "dpp_subsample_fit_pca.py")
    spec = importlib.util.spec_from_file_location("example_workflow", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    r = mod.main(n=300, natoms=8, K=12, batch=40, verbose=False)
    assert len(set(r["subset"])) == 40 and abs(r["inclusion_sum"] - 40.0) < 1e-6
    assert r["beta_err"] < 1e-2 and r["e_mae"] < 1e-2 and r["f_mae"] < 1e-2
    assert r["pca_dim"] == 8 and 0.0 < r["var_kept"] <= 1.0


def test_device_kdpp_sampler(pl, orc):
    """pl_ell_sample: projection-DPP sampling over selected eigenvectors of the resident ensemble. (a) With the same eigenvectors and
    the same uniforms it picks exactly the items of the numpy restatement of the algorithm; (b) the empirical inclusion frequencies of
    many k-DPP draws match the exact inclusion probabilities of the k-DPP... of the L-ensemble rescaled to expected size k, within
    sampling error; (c) a draw has k distinct items."""
    import ctypes

    from plb200 import _lib

    rng = np.random.default_rng(101)
    n, d = 300, 10
    X = rng.standard_normal((n, d)) * 0.6 + rng.standard_normal(d)
    ell = pl.EllEnsemble.from_features(X, pl.RBF(pl.Euclidean(d), ell=2.0))
    i64p = ctypes.POINTER(ctypes.c_int64)
    for m in (1, 7, 40):
        J = np.sort(rng.choice(np.arange(n - 80, n), size=m, replace=False)).astype(np.int64)  # among the 80 largest eigenvalues
        u = rng.random(m)
        out = np.empty(m, dtype=np.int64)
        V = ell.eigvecs(J)  # the same device-resident eigenvectors, downloaded
        ref = orc.sample_projection_dpp(V, u)
        for form in ("kernel-row", "gram-schmidt"):  # n <= 32768 takes the kernel-row (Cholesky) form; PLB200_DPP_GS forces the other
            if form == "gram-schmidt":
                os.environ["PLB200_DPP_GS"] = "1"
            try:
                out[:] = -7
                _lib.check(_lib.lib().pl_ell_sample(J.ctypes.data_as(i64p), m, n, _lib.dptr(u), out.ctypes.data_as(i64p)), "pl_ell_sample")
            finally:
                os.environ.pop("PLB200_DPP_GS", None)
            assert list(out) == ref, (form, m, list(out), ref)
            assert len(set(out.tolist())) == m
    # statistics of full k-DPP draws through the mirror (phase 1 on the host, phase 2 on the device)
    n2, k = 40, 6
    X2 = rng.standard_normal((n2, 5)) + 1.0
    ell2 = pl.EllEnsemble.from_features(X2, pl.RBF(pl.Euclidean(5), ell=1.5))
    assert ell2._resident
    r = np.random.default_rng(5)
    ndraw = 3000
    cnt = np.zeros(n2)
    for _ in range(ndraw):
        s = pl.sample(ell2, k, rng=r)
        assert len(s) == k and len(set(s)) == k
        cnt[s] += 1
    # exact k-DPP inclusion probabilities: P(i in S) = sum_j U_ij^2 * lam_j e_{k-1}(lam without j) / e_k(lam)
    lam = ell2.λ
    U = ell2.eigvecs(np.arange(n2))

    def esp(vals, kk):
        e = np.zeros(kk + 1)
        e[0] = 1.0
        for x in vals:
            e[1:] = e[1:] + x * e[:-1]
        return e[kk]

    ek = esp(lam, k)
    w = np.array([lam[j] * esp(np.delete(lam, j), k - 1) / ek for j in range(n2)])
    pin = (U * U) @ w
    assert abs(pin.sum() - k) < 1e-8
    se = np.sqrt(pin * (1 - pin) / ndraw)
    assert np.all(np.abs(cnt / ndraw - pin) < 5 * se + 1e-3), float(np.max(np.abs(cnt / ndraw - pin) / se))


def test_device_esp_selection(pl, orc):
    """pl_esp_select (first phase of a k-DPP draw, Kulesza & Taskar Alg. 8, built by one thread-block cluster): the eigenvectors of the
    oracle's restatement for the same eigenvalues and uniforms — spread-out spectra (hundreds of orders of magnitude in e_k), zero
    eigenvalues, k = 1, k = n - 1, k = number of positive eigenvalues; through the mirror, a draw with the table on the device equals the
    draw with the table on the host for the same generator."""
    import ctypes

    from plb200 import _lib

    rng = np.random.default_rng(404)
    i64p = ctypes.POINTER(ctypes.c_int64)

    def dev(lam, k, u):
        idx = np.full(max(k, 1), -5, dtype=np.int64)
        cnt = ctypes.c_int64(-1)
        _lib.check(_lib.lib().pl_esp_select(_lib.dptr(lam), len(lam), k, _lib.dptr(u), idx.ctypes.data_as(i64p), ctypes.byref(cnt)), "pl_esp_select")
        return [int(i) for i in idx[: cnt.value]]

    cases = []
    for N, k in [(40, 6), (257, 31), (300, 1), (64, 63), (1000, 400), (9000, 20)]:
        cases.append((np.sort(rng.random(N) ** 4 * 50.0), k))
    lamz = np.sort(rng.random(120) * 2.0)
    lamz[:70] = 0.0
    cases += [(lamz, 50), (lamz, 17), (np.sort(10.0 ** rng.uniform(-30, 30, 500)), 120)]
    for lam, k in cases:
        lam = np.ascontiguousarray(lam)
        u = rng.random(len(lam))
        got = dev(lam, k, u)
        ref = orc.select_eigenvectors_kdpp(lam, k, u)
        assert got == ref, (len(lam), k, got[:10], ref[:10])
        assert len(got) == k
    assert dev(lamz, 0, rng.random(120)) == []
    with pytest.raises(_lib.PLB200Error):
        dev(lamz, 51, rng.random(120))  # only 50 positive eigenvalues
    with pytest.raises(_lib.PLB200Error):
        dev(np.array([1.0, -1.0, 2.0]), 1, rng.random(3))
    # through the mirror
    X = rng.standard_normal((200, 6)) * 0.5 + 1.0
    ell = pl.EllEnsemble.from_features(X, pl.RBF(pl.Euclidean(6), ell=1.2))
    pl.rescale_(ell, 30)
    a = pl.sample(ell, 30, rng=np.random.default_rng(9))
    os.environ["PLB200_ESP_HOST"] = "1"
    try:
        b = pl.sample(ell, 30, rng=np.random.default_rng(9))
    finally:
        os.environ.pop("PLB200_ESP_HOST", None)
    assert a == b and len(set(a)) == 30


def test_device_greedy_map(pl, orc):
    """pl_ell_greedy (get_dpp_mode, dpp.jl:59-62): the device greedy MAP picks the items of the numpy restatement on the same kernel
    matrix, in the same order; a rank-deficient ensemble stops at its numerical rank."""
    rng = np.random.default_rng(55)
    n, d = 500, 9
    X = rng.standard_normal((n, d)) * 0.7 + rng.standard_normal(d)
    for k in (1, 5, 60):
        dpp = pl.kDPP(X, pl.RBF(pl.Euclidean(d), ell=1.2), batch_size=k)
        mode = pl.get_dpp_mode(dpp)
        ref = orc.greedy_map(dpp.K.L, k)
        assert mode == ref and len(set(mode)) == k
    # DotProduct(alpha = 1) of 3-dimensional features has rank 3... (x.y/|x||y|): the greedy stops when the variances are exhausted
    X3 = rng.standard_normal((200, 3)) + 2.0
    ell = pl.EllEnsemble.from_features(X3, pl.DotProduct(1))
    g = pl.greedy_subset(ell, 10)
    assert 3 <= len(g) <= 10 and len(set(g)) == len(g)
    assert g[:3] == orc.greedy_map(ell.L, 3)


def test_dbscan_distance_matrices(pl, orc):
    """distance_matrix_periodic / distance_matrix_kabsch (dbscan.jl:132-170, kabsch.jl) on the GPU against the literal restatements:
    the periodic form is bit-identical (same operations in the same order), the Kabsch form agrees to rounding with the LAPACK-SVD
    evaluation, is invariant under rigid motion of either configuration, and handles planar configurations and reflections."""
    rng = np.random.default_rng(77)
    n, natoms = 23, 17
    box = np.array([7.0, 8.5, 6.25])
    base = rng.uniform(0, 1, size=(natoms, 3)) * box
    pos = np.stack([np.mod(base + 0.4 * rng.standard_normal((natoms, 3)) + rng.uniform(-3, 3, size=3), box) for _ in range(n)])
    D = pl.distance_matrix_periodic(pos, box)
    Do = orc.distance_matrix(pos, box)
    assert np.array_equal(D, Do) and np.array_equal(D, D.T) and np.all(np.diag(D) == 0.0)
    # Kabsch: configurations = rigidly moved, noisy copies of two shapes
    shapes = [rng.standard_normal((natoms, 3)) * 2.0, rng.standard_normal((natoms, 3)) * 2.0]

    def rigid(X):
        Q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        if np.linalg.det(Q) < 0:
            Q[:, 0] = -Q[:, 0]
        return X @ Q + rng.uniform(-5, 5, size=3)

    pk = np.stack([rigid(shapes[i % 2] + 0.05 * rng.standard_normal((natoms, 3))) for i in range(n)])
    Dk = pl.distance_matrix_kabsch(pk)
    Dko = orc.distance_matrix(pk)
    assert np.max(np.abs(Dk - Dko)) < 1e-10 * np.max(Dko) and np.array_equal(Dk, Dk.T) and np.all(np.diag(Dk) == 0.0)
    assert Dk[0, 2] < 0.3 and Dk[0, 1] > 1.0  # same shape: only the noise is left; different shapes stay apart
    pk2 = pk.copy()
    pk2[5] = rigid(pk[5])
    assert np.max(np.abs(pl.distance_matrix_kabsch(pk2) - Dk)) < 1e-9
    # a mirror image needs the reflection guard (f = -1): the RMSD must NOT vanish; planar configurations (rank-2 covariance) work
    mirrored = shapes[0] * np.array([1.0, 1.0, -1.0])
    pair = np.stack([shapes[0], mirrored])
    assert abs(pl.distance_matrix_kabsch(pair)[0, 1] - orc.kabsch_rmsd(pair[0], pair[1])) < 1e-10 and pl.distance_matrix_kabsch(pair)[0, 1] > 0.1
    planar = rng.standard_normal((2, natoms, 3))
    planar[:, :, 2] = 0.0
    assert abs(pl.distance_matrix_kabsch(planar)[0, 1] - orc.kabsch_rmsd(planar[0], planar[1])) < 1e-10


def test_reference_dbscan_selector_tests(pl):
    """subset_selector_tests.jl:28-39 re-stated on synthetic configurations (the reference loads an xyz file): DBSCANSelector over the
    GPU-built distance matrix finds the planted clusters, get_random_subset draws sample_size ÷ nclusters indices from each."""
    rng = np.random.default_rng(8)
    natoms, box = 12, np.array([9.0, 9.0, 9.0])
    centres = [rng.uniform(0, 9, size=(natoms, 3)) for _ in range(3)]
    pos = np.stack([np.mod(centres[i % 3] + 0.01 * rng.standard_normal((natoms, 3)), box) for i in range(45)])
    sel = pl.DBSCANSelector(pos, 0.05, 5, 6, box_lengths=box)
    assert isinstance(sel, pl.SubsetSelector)
    assert len(sel.clusters) == 3 and sorted(len(c) for c in sel.clusters) == [15, 15, 15]
    for c in sel.clusters:
        assert len({i % 3 for i in c}) == 1
    sub = pl.get_random_subset(sel, rng=np.random.default_rng(0))
    assert len(sub) == 6 and all(isinstance(i, int) for i in sub)
    assert sorted(sum(1 for i in sub if i in c) for c in sel.clusters) == [2, 2, 2]
    # non-periodic: Kabsch RMSD separates rigidly moved copies of different shapes
    shapes = [rng.standard_normal((natoms, 3)) * 2 for _ in range(2)]
    pk = []
    for i in range(30):
        Q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        pk.append((shapes[i % 2] + 0.005 * rng.standard_normal((natoms, 3))) @ Q + rng.uniform(-4, 4, size=3))
    sel2 = pl.DBSCANSelector(np.stack(pk), 0.1, 5, 4)
    assert len(sel2.clusters) == 2 and all(len({i % 2 for i in c}) == 1 for c in sel2.clusters)


def test_metrics(pl):
    """mae / rmse / rsq / mean_cos / get_metrics (src/Metrics/metrics.jl:11-75) against the definitions evaluated in numpy."""
    rng = np.random.default_rng(6)
    for n in (3, 999, 300_000):
        x = rng.standard_normal(n) * 2.0 + 5.0
        xp = x + 0.1 * rng.standard_normal(n)
        if n >= 999:
            x[3:6] = 0.0  # a zero force vector: its cosine is NaN and must be filtered (metrics.jl:51)
        d = xp - x
        ref = {"mae": np.sum(np.abs(d)) / n, "rmse": np.sqrt(np.sum(d ** 2) / n), "rsq": 1 - np.sum(d ** 2) / np.sum((x - np.mean(x)) ** 2)}
        xv, pv = x.reshape(-1, 3), xp.reshape(-1, 3)
        with np.errstate(all="ignore"):
            cs = np.sum(xv * pv, axis=1) / (np.linalg.norm(xv, axis=1) * np.linalg.norm(pv, axis=1))
        ref["mean_cos"] = np.mean(cs[~np.isnan(cs)])
        assert abs(pl.mae(xp, x) - ref["mae"]) < 1e-12 * ref["mae"]
        assert abs(pl.rmse(xp, x) - ref["rmse"]) < 1e-12 * ref["rmse"]
        assert abs(pl.rsq(xp, x) - ref["rsq"]) < 1e-12
        assert abs(pl.mean_cos(xp, x) - ref["mean_cos"]) < 1e-12
        m = pl.get_metrics(xp, x, metrics=[pl.mae, pl.rmse, pl.rsq], label="e")
        assert list(m) == ["e_mae", "e_rmse", "e_rsq"] and abs(m["e_rmse"] - ref["rmse"]) < 1e-12 * ref["rmse"]
    assert np.isnan(pl.mean_cos(np.ones(4), np.ones(4)))  # not a multiple of 3: no 3-vectors
    with pytest.raises(ValueError):
        pl.mae(np.ones(3), np.ones(4))


def test_pageable_staging_paths(pl, orc):
    """Pageable host memory goes through the library's pinned staging ring + host threads (hoststage.h): same results as the direct
    path (PLB200_NO_STAGING=1), for the streamed kernel-matrix download, the pitched matrix upload and the block-list upload."""
    import os

    rng = np.random.default_rng(12)
    X = rng.standard_normal((2500, 40)) * 0.5 + 1.0  # 2500^2 * 8 = 50 MB result: staged (>= 4 MB), T = 20 >= 16: the streamed path
    k = pl.RBF(pl.Euclidean(40), ell=2.0)
    K1 = np.asarray(pl.KernelMatrix(X, k, fill="U"))
    K1f = np.asarray(pl.KernelMatrix(X, k))
    Kx1 = pl.KernelMatrix(X, X[:700], k)  # rectangular 2500 x 700 = 14 MB
    A = rng.standard_normal((120_000, 37))  # 35 MB, odd K: pitched upload
    b = rng.standard_normal(120_000)
    M1, v1 = pl.normal_equations(A, b, None, 500, 30.0, 1.0, True, 0.01)
    blocks = [rng.standard_normal(64) for _ in range(50)] + [np.asfortranarray(rng.standard_normal((64, 3 * n))) for n in rng.integers(20, 200, size=150)]
    bb = rng.standard_normal(sum(1 if m.ndim == 1 else m.shape[1] for m in blocks))
    M3, v3 = pl.normal_equations_blocks(blocks, bb, None, 50, 100.0, 1.0, True, 0.0, chunk_rows=4096)
    os.environ["PLB200_NO_STAGING"] = "1"
    try:
        K2 = np.asarray(pl.KernelMatrix(X, k, fill="U"))
        K2f = np.asarray(pl.KernelMatrix(X, k))
        Kx2 = pl.KernelMatrix(X, X[:700], k)
        M2, v2 = pl.normal_equations(A, b, None, 500, 30.0, 1.0, True, 0.01)
        M4, v4 = pl.normal_equations_blocks(blocks, bb, None, 50, 100.0, 1.0, True, 0.0, chunk_rows=4096)
    finally:
        del os.environ["PLB200_NO_STAGING"]
    assert np.array_equal(K1, K2) and np.array_equal(K1f, K2f) and np.array_equal(Kx1, Kx2)
    assert np.array_equal(M1, M2) and np.array_equal(v1, v2) and np.array_equal(M3, M4) and np.array_equal(v3, v4)
    assert relerr(K1f[:200, :200], orc.symmetric(orc.kernel_matrix(X[:200], orc.RBF(orc.Euclidean(40), ell=2.0)))) < RTOL


def test_workspace_trim(pl, orc):
    """pl_workspace_trim: scratch goes back to the driver; with keep_resident the resident L-ensemble keeps answering (same inclusion
    probabilities, same draw), without it the pl_ell_* entries ask for a new ensemble; everything regrows on demand."""
    rng = np.random.default_rng(77)
    X = rng.standard_normal((700, 12)) * 0.5 + 1.0
    k = pl.RBF(pl.Euclidean(12), ell=1.5)
    ell = pl.EllEnsemble.from_features(X, k)
    pl.rescale_(ell, 60)
    p0 = pl.inclusion_prob(ell)
    s0 = pl.sample(ell, 60, rng=np.random.default_rng(3))
    before = pl.workspace_bytes()
    pl.workspace_trim(True)
    mid = pl.workspace_bytes()
    assert 0 < mid < before
    assert np.array_equal(pl.inclusion_prob(ell), p0)
    assert pl.sample(ell, 60, rng=np.random.default_rng(3)) == s0
    K1 = np.asarray(pl.KernelMatrix(list(X), k))
    pl.workspace_trim(False)
    assert pl.workspace_bytes() == 0
    assert not ell._resident
    import ctypes

    from plb200 import _lib

    out = np.empty(700)
    assert _lib.lib().pl_ell_inclusion_prob(1.0, 700, _lib.dptr(out)) != 0  # no resident ensemble any more
    assert np.allclose(pl.inclusion_prob(ell), p0, rtol=0, atol=1e-12)  # the mirror downloaded U before the eviction (n <= 4096)
    assert np.array_equal(np.asarray(pl.KernelMatrix(list(X), k)), K1)


def test_syrk_schedule_modes_and_streams(pl, orc):
    """The SYRK work table is cached on the device per shape: a second stream using the cached table, a shape change and a trim in
    between all give the bits of the first result (fixed-order second stage), and the result matches the oracle."""
    import torch

    from plb200.device import DeviceEngine

    eng = DeviceEngine()
    rng = np.random.default_rng(31)
    R, K, ne = 5000, 200, 40
    A = rng.standard_normal((R, K))
    b = rng.standard_normal(R)
    Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    S0, v0 = eng.syrk_partial(Ad, None, ne, 3.0, 0.5, bd, None, True)
    torch.cuda.synchronize()
    s2 = torch.cuda.Stream()
    with torch.cuda.stream(s2):
        S1, v1 = eng.syrk_partial(Ad, None, ne, 3.0, 0.5, bd, None, True)  # cached table, other stream
    s2.synchronize()
    assert torch.equal(S0, S1) and torch.equal(v0, v1)
    eng.syrk_partial(Ad[:3333], None, ne, 3.0, 0.5, bd[:3333], None, True)  # another shape replaces the table
    pl.workspace_trim(True)
    S2, v2 = eng.syrk_partial(Ad, None, ne, 3.0, 0.5, bd, None, True)
    torch.cuda.synchronize()
    assert torch.equal(S0, S2) and torch.equal(v0, v2)
    w = np.where(np.arange(R) < ne, 3.0, 0.5)
    Sref = (A * w[:, None]).T @ A
    assert normerr(S0.cpu().numpy(), Sref) < RTOL
    assert np.allclose(v0.cpu().numpy()[:K], A.T @ (w * b), rtol=0, atol=1e-9 * np.abs(A.T @ (w * b)).max())


def test_gram_dev_graph_replay_tracks_buffer_contents(pl, orc, monkeypatch):
    """pl_gram_dev replays a cached CUDA graph from the third identical call on (same pointers / shapes / parameters): the replay must read the
    buffers as they are NOW, survive a workspace trim (new workspace pointers -> new graph), count its launches, and equal the plain launches
    bit for bit; both tile shapes are exercised."""
    import torch

    from plb200 import _lib

    lib = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(9)
    for n, d, kid, k_or in ((700, 26, _lib.PL_KERNEL_DOT, orc.DotProduct()), (2300, 40, _lib.PL_KERNEL_RBF, orc.RBF(orc.Euclidean(40)))):
        X = torch.empty((n, d), device="cuda", dtype=torch.float64)
        K = torch.zeros((n, n), device="cuda", dtype=torch.float64)
        st = torch.cuda.current_stream().cuda_stream

        def call():
            _lib.check(lib.pl_gram_dev(kid, X.data_ptr(), n, d, None, 0, 0, d, 2, 0, None, 1.0, 1.0, K.data_ptr(), n, _lib.PL_FILL_FULL, 0, 1, 0, st), "pl_gram_dev")

        per_call = None
        for it in range(6):
            X.copy_(torch.randn((n, d), device="cuda", dtype=torch.float64, generator=g) * 0.3 + 1.0 + it)
            if it == 4:
                _lib.check(lib.pl_workspace_trim(1), "trim")  # workspace pointers change: the cached graph must not be replayed
            l0 = pl.launch_count()
            call()
            torch.cuda.synchronize()
            nl = pl.launch_count() - l0
            per_call = per_call or nl
            assert nl == per_call, "a replay accounts for the same number of kernel launches"
            got = K.cpu().numpy()
            monkeypatch.setenv("PLB200_NO_GRAPH", "1")
            K.zero_()
            call()
            torch.cuda.synchronize()
            monkeypatch.delenv("PLB200_NO_GRAPH")
            assert np.array_equal(got, K.cpu().numpy()), (n, it)
            if it in (0, 5):
                rows = [0, n // 2, n - 1]
                Xh = X.cpu().numpy()
                ref = orc.kernel_matrix_cross(Xh[rows], Xh, k_or)
                gr = got[rows]
                for q, i in enumerate(rows):
                    gr[q, i] = ref[q, i]  # the diagonal is forced (k(x,x)), the oracle rounds it
                assert relerr(gr, ref) < RTOL


def test_tile_shapes_agree_bitwise(pl, monkeypatch):
    """64 x 64 and 128 x 128 tiles run the same contraction order and the same epilogue arithmetic: identical bits, symmetric and rectangular,
    with ragged edges."""
    rng = np.random.default_rng(91)
    for n, m, d in ((333, 0, 26), (1000, 0, 7), (257, 190, 33), (129, 640, 8)):
        X = rng.standard_normal((n, d)) * 0.3 + 1.0
        Y = rng.standard_normal((m, d)) * 0.3 + 1.0 if m else None
        for k in (pl.DotProduct(), pl.RBF(pl.Euclidean(d), ell=1.3, beta=0.8), pl.InverseMultiquadric(pl.Euclidean(d))):
            res = []
            for small in ("0", "1"):
                monkeypatch.setenv("PLB200_GRAM_SMALL", small)
                res.append(np.asarray(pl.KernelMatrix(X, k)) if Y is None else pl.KernelMatrix(X, Y, k))
            monkeypatch.delenv("PLB200_GRAM_SMALL")
            assert np.array_equal(res[0], res[1]), (n, m, d, type(k).__name__)


def test_packed_shard_tail_wave_subtiles(pl, orc, monkeypatch):
    """Tile-packed shards of a few waves finish their last PARTIAL wave as 64 x 64 sub-tiles written into the same packed tiles (TAIL mode of the
    small-tile kernel). Same bits as the plain 128 x 128 launch, for every part, with ragged edges, odd tile counts and diagonal parents (whose
    (1, 0) sub-tile is the mirror image of (0, 1), never computed on its own); sampled tiles against the oracle."""
    import ctypes

    import torch

    from plb200 import _lib

    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(17)
    engaged = 0
    for n, d, nparts in ((2700, 26, 1), (2689, 40, 1), (4000, 17, 2), (3777, 64, 2), (5400, 12, 3)):
        X = torch.randn((n, d), device="cuda", dtype=torch.float64, generator=g) * 0.3 + 1.0
        Xh = X.cpu().numpy()
        for kid, ko in ((_lib.PL_KERNEL_DOT, orc.DotProduct()), (_lib.PL_KERNEL_RBF, orc.RBF(orc.Euclidean(d)))):
            for part in range(nparts):
                nt = int(lib.pl_gram_tile_count(n, 0, part, nparts))
                r = nt % 148
                engaged += int(nt >= 148 and 0 < 4 * r <= 444)
                outs = []
                for no_tail in ("1", None):
                    if no_tail:
                        monkeypatch.setenv("PLB200_GRAM_NO_TAIL", no_tail)
                    else:
                        monkeypatch.delenv("PLB200_GRAM_NO_TAIL")
                    tiles = torch.full((nt, 128, 128), -3.0, device="cuda", dtype=torch.float64)
                    _lib.check(lib.pl_gram_dev(kid, X.data_ptr(), n, d, None, 0, 0, d, 2, 0, None, 1.0, 1.0, tiles.data_ptr(), 128, _lib.PL_FILL_FULL, part, nparts, 1, st),
                               "pl_gram_dev")
                    torch.cuda.synchronize()
                    outs.append(tiles.cpu().numpy())
                assert np.array_equal(outs[0], outs[1]), (n, d, nparts, part, kid)
                ti, tj = ctypes.c_int64(), ctypes.c_int64()
                for k in (nt - 1, nt - 2, nt - r, max(nt - r - 1, 0), nt - (r + 1) // 2):  # tail tiles and the last main tile
                    _lib.check(lib.pl_gram_tile_coords(n, 0, part, nparts, int(k), ctypes.byref(ti), ctypes.byref(tj)), "coords")
                    if ti.value < 0:
                        continue
                    r0, r1, c0, c1 = ti.value * 128, min(ti.value * 128 + 128, n), tj.value * 128, min(tj.value * 128 + 128, n)
                    ref = orc.kernel_matrix_cross(Xh[r0:r1], Xh[c0:c1], ko)
                    got = outs[1][k, : r1 - r0, : c1 - c0]
                    if ti.value == tj.value:
                        np.fill_diagonal(ref, 1.0)
                        assert np.array_equal(got, got.T), "packed diagonal tile: both triangles, exactly symmetric"
                    assert relerr(got, ref) < RTOL, (n, part, k)
    assert engaged >= 6, "the shapes above must actually take the tail path"
