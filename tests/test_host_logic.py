"""Host-side logic of the reference-interface mirror that needs no GPU: containers, LinearProblem layout and packing,
Symmetric storage semantics, row sharding, the restated k-DPP routines, eigen-direction selection."""
import numpy as np
import pytest

import plb200 as pl
from oracle import oracle as orc
from plb200.dimred import _select
from plb200.distributed import shard_rows


def make_ds(rng, ncfg, nat, d, energies=True, forces=True):
    cfgs = []
    for _ in range(ncfg):
        items = [pl.LocalDescriptors(rng.standard_normal((nat, d)))]
        if energies:
            items.append(pl.Energy(rng.random()))
        if forces:
            items += [pl.ForceDescriptors(rng.standard_normal((nat, 3, d))), pl.Forces(rng.standard_normal((nat, 3)))]
        cfgs.append(pl.Configuration(*items))
    return pl.DataSet(cfgs)


def test_features_match_reference_definition():
    rng = np.random.default_rng(0)
    ds = make_ds(rng, 6, 20, 8, energies=False, forces=False)
    gm, gs = pl.GlobalMean(), pl.GlobalSum()
    for c in ds:
        b = pl.get_values(pl.get_local_descriptors(c))
        assert np.array_equal(pl.compute_feature(c, gm), orc.global_mean(list(b)))
        assert np.array_equal(pl.compute_feature(c, gs), orc.global_sum(list(b)))
        assert np.array_equal(pl.compute_feature(c, gm, dt=pl.ForceDescriptors), pl.compute_feature(c, gm))  # dt ignored (features.jl:59-61)
    # compute_features over a whole DataSet is a GPU segmented reduction (pl_global_feature): see tests/test_gpu_parity.py
    L, off = pl.stack_locals([pl.get_values(pl.get_local_descriptors(c)) for c in ds])
    assert L.shape == (120, 8) and list(off) == [0, 20, 40, 60, 80, 100, 120]
    with pytest.raises(KeyError):
        pl.get_energy(ds[0])


def test_linear_problem_layout_matches_oracle():
    rng = np.random.default_rng(1)
    ds = make_ds(rng, 7, 5, 4)
    lp = pl.LinearProblem(ds)
    assert isinstance(lp, pl.CovariateLinearProblem) and isinstance(lp, pl.LinearProblem)
    o = orc.linear_problem(
        [list(pl.get_values(pl.get_local_descriptors(c))) for c in ds], [pl.get_values(pl.get_energy(c)) for c in ds],
        [[list(a) for a in pl.get_values(pl.get_force_descriptors(c))] for c in ds], [list(pl.get_values(pl.get_forces(c))) for c in ds])
    for a, b in zip(lp.B, o.B):
        assert np.array_equal(a, b)
    for a, b in zip(lp.dB, o.dB):
        assert a.shape == b.shape and np.array_equal(a, b)
    for a, b in zip(lp.f, o.f):
        assert np.array_equal(a, b)
    A, W, b = pl.assemble_matrices(lp, [100.0, 1.0])
    Ao, bo, qo = orc.wls_matrices(o, [100.0, 1.0], False)
    assert np.array_equal(A, Ao) and np.array_equal(b, bo) and np.array_equal(W, qo)
    assert A.flags.c_contiguous  # the device layout: row-major R x K, zero transposition from Julia's column-major blocks
    # energies only / forces only -> Univariate
    lpe = pl.LinearProblem(make_ds(rng, 4, 5, 4, forces=False))
    assert isinstance(lpe, pl.UnivariateLinearProblem) and pl.pack_rows(lpe.iv_data).shape == (4, 4)
    ds_f = pl.DataSet([pl.Configuration(pl.ForceDescriptors(rng.standard_normal((5, 3, 4))), pl.Forces(rng.standard_normal((5, 3)))) for _ in range(3)])
    lpf = pl.LinearProblem(ds_f)
    assert isinstance(lpf, pl.UnivariateLinearProblem) and lpf.iv_data[0].shape == (4, 15) and pl.pack_rows(lpf.iv_data).shape == (45, 4)
    with pytest.raises(ValueError):
        pl.LinearProblem(pl.DataSet([pl.Configuration(pl.Energy(1.0))]))
    bad = pl.DataSet([pl.Configuration(pl.LocalDescriptors(np.ones((2, 3))), pl.Energy(1.0), pl.ForceDescriptors(np.ones((2, 3, 4))), pl.Forces(np.ones((2, 3))))])
    with pytest.raises(ValueError, match="different dimension"):
        pl.LinearProblem(bad)


def test_symmetric_wrapper():
    d = np.array([[1.0, 2.0, 3.0], [0.0, 4.0, 5.0], [0.0, 0.0, 6.0]])
    S = pl.Symmetric(d)
    assert np.array_equal(np.asarray(S), np.array([[1, 2, 3], [2, 4, 5], [3, 5, 6.0]]))
    assert S[2, 0] == 3.0 and S[0, 2] == 3.0 and S.shape == (3, 3)
    with pytest.raises(ValueError):
        pl.Symmetric(np.ones((2, 3)))


def test_kernel_argument_errors():
    with pytest.raises(TypeError):
        pl.RBF("forstner")
    with pytest.raises(TypeError):
        pl.DotProduct(2.0)
    assert pl.get_parameters(pl.DotProduct(3)) == (3,)
    assert pl.get_parameters(pl.RBF(pl.Euclidean(4), ell=2.0)) == (1e-8, 2.0, 1.0)
    e = pl.Euclidean(pl.Symmetric(np.eye(3)))
    assert e.dim == 3 and e.kind == 2
    assert pl.Euclidean(pl.Diagonal([1.0, 2.0])).kind == 1


def test_shard_rows_partitions_exactly():
    for n, world, align in ((10, 3, 1), (2890000, 8, 289), (7, 8, 1), (1000, 4, 96)):
        spans = [shard_rows(n, r, world, align) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0 and a0 <= a1
        assert all(lo % align == 0 for lo, _ in spans)
        sizes = [(hi - lo + align - 1) // align for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_select_eigendirections_rules():
    lam = np.array([0.1, 0.4, 1.5, 8.0])  # ascending, as eigen() returns
    phi = np.eye(4)
    l2, W = _select(lam, phi, 2)
    assert np.array_equal(l2, lam[::-1]) and W.shape == (4, 2) and W[3, 0] == 1.0
    l3, W = _select(lam, phi, 0.1)
    resid = 1.0 - np.cumsum(lam[::-1]) / lam.sum()
    assert W.shape[1] == int(np.sum(resid > 0.1))


def test_kdpp_host_routines_on_oracle_kernel():
    """The restated Determinantal.jl pieces (host-side, [external]) on a CPU-built L-ensemble."""
    rng = np.random.default_rng(5)
    X = rng.standard_normal((40, 6)) + 1.0
    K = orc.symmetric(orc.kernel_matrix(X, orc.DotProduct()))
    ell = pl.EllEnsemble.from_eigen(*np.linalg.eigh(K), L=K)  # the GPU eigendecomposition (pl_eigh) is covered by the gpu tests
    pl.rescale_(ell, 5)
    ip = pl.inclusion_prob(ell)
    assert abs(ip.sum() - 5.0) < 1e-8 and np.all(ip > 0) and np.all(ip < 1)
    s = pl.sample(ell, 5, rng=np.random.default_rng(0))
    assert len(s) == 5 and len(set(s)) == 5 and all(0 <= i < 40 for i in s)
    assert s == pl.sample(ell, 5, rng=np.random.default_rng(0)), "same seed + same eigendecomposition -> same indices"
    g = pl.greedy_subset(ell, 5)
    assert len(set(g)) == 5
    # empirical inclusion frequencies of the k-DPP sampler are in the right ballpark of a k-DPP (sum to k)
    cnt = np.zeros(40)
    r = np.random.default_rng(1)
    for _ in range(300):
        for i in pl.sample(ell, 5, rng=r):
            cnt[i] += 1
    assert abs(cnt.sum() / 300 - 5.0) < 1e-12
    with pytest.raises(ValueError):
        pl.rescale_(ell, 30)  # a degree-2 polynomial kernel of 6-dim features has rank 21 < 30


def test_random_selector():
    """RandomSelector / get_random_subset (random.jl): a prefix of a random permutation."""
    r = pl.RandomSelector(10, batch_size=2)
    assert isinstance(r, pl.SubsetSelector)
    s = pl.get_random_subset(r, rng=np.random.default_rng(0))
    assert len(s) == 2 and len(set(s)) == 2 and all(isinstance(i, int) and 0 <= i < 10 for i in s)
    assert sorted(pl.get_random_subset(pl.RandomSelector(7), rng=np.random.default_rng(1))) == list(range(7))


def test_kdpp_eigenvector_selection_host_vs_oracle():
    """phase 1 of a k-DPP draw (Kulesza & Taskar Alg. 8): the mirror's numpy form picks the oracle's eigenvectors for the same uniforms,
    and the oracle's selection frequencies match the exact probabilities lam_j e_{k-1}(lam \\ j) / e_k(lam)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle as orc
    from plb200 import dpp

    rng = np.random.default_rng(0)
    for N, k in [(40, 6), (200, 50), (300, 1), (50, 50), (60, 59)]:
        lam = np.sort(rng.random(N) ** 3 * 5)
        u = rng.random(N)
        a = orc.select_eigenvectors_kdpp(lam, k, u)
        assert a == dpp._esp_select_host(lam, k, u) and len(a) == k and sorted(a, reverse=True) == a
    N, k, nd = 12, 4, 6000
    lam = rng.random(N) * 3

    def esp(v, kk):
        e = np.zeros(kk + 1)
        e[0] = 1
        for x in v:
            e[1:] = e[1:] + x * e[:-1]
        return e[kk]

    w = np.array([lam[j] * esp(np.delete(lam, j), k - 1) / esp(lam, k) for j in range(N)])
    cnt = np.zeros(N)
    for _ in range(nd):
        cnt[orc.select_eigenvectors_kdpp(lam, k, rng.random(N))] += 1
    assert np.max(np.abs(cnt / nd - w) / np.sqrt(w * (1 - w) / nd)) < 5.0
