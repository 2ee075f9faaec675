# Julia-side suite of the drop-in (needs Julia >= 1.10, PotentialLearning.jl v0.2.7 and a B200; NOT runnable in the repository's build
# image, which has no Julia — tests/ drives the same C ABI from Python and from plain C instead).
# Re-states the hot-path assertions of the reference's own tests (test/kernels/kernel_tests.jl, test/learning/linear_tests.jl,
# test/dimension_reduction/dimension_reduction_tests.jl) and adds value-level parity at 1e-10 against the reference's generic methods,
# reached with `invoke` (more specific overrides) or re-stated from the reference's formulas (overwritten methods).
using Test, Random, LinearAlgebra, Statistics
using PotentialLearning
using PotentialLearningB200
const B = PotentialLearningB200

relerr(a, b) = maximum(abs.(a .- b) ./ abs.(b))
const RefF = Union{Vector{Vector{Float64}},Vector{Symmetric{Float64,Matrix{Float64}}}}

@testset "PotentialLearningB200" begin
    @test B.ACTIVE[]
    Random.seed!(7)
    d, num_atoms, num_configs = 8, 20, 10                                              # kernel_tests.jl:7-9
    ld = [LocalDescriptors([randn(d) for _ = 1:num_atoms]) for _ = 1:num_configs]
    ds = DataSet(Configuration.(ld))
    gm = GlobalMean()
    f_gm = compute_feature.(ld, (gm,))
    @test compute_features(ds, gm) == f_gm                                              # kernel_tests.jl:32 — bit-identical (Julia's summation order)
    dp, rbf = DotProduct(), RBF(Euclidean(d))
    for k in (dp, rbf)
        K = KernelMatrix(f_gm, k)
        @test typeof(K) <: Symmetric{Float64,Matrix{Float64}}                           # kernel_tests.jl:85,87
        @test K.uplo == 'U' && all(iszero, tril(K.data, -1))                            # zeros(n,n) storage, :U triangle only
        Kref = invoke(KernelMatrix, Tuple{RefF,PotentialLearning.Kernel}, f_gm, k)
        @test relerr(Matrix(K), Matrix(Kref)) < 1e-10
        Kx = KernelMatrix(f_gm[1:4], f_gm[5:10], k)
        @test Kx isa Matrix{Float64} && size(Kx) == (4, 6)
        @test relerr(Kx, invoke(KernelMatrix, Tuple{RefF,RefF,PotentialLearning.Kernel}, f_gm[1:4], f_gm[5:10], k)) < 1e-10
    end
    @test all(KernelMatrix(f_gm, rbf)[i, i] == 1.0 for i = 1:num_configs)               # K(x,x) = β exactly; RBF.α never applied
    dpp = kDPP(ds, gm, dp; batch_size = 2)                                              # subset_selector_tests.jl:17-21
    @test get_random_subset(dpp) isa Vector{<:Int}

    # learn!: WLS against the reference's own expression (linear-learn.jl:232-253)
    Random.seed!(0)
    N, natoms, K = 30, 5, 8
    lp = PotentialLearning.CovariateLinearProblem(randn(N), [randn(3natoms) for _ = 1:N], [randn(K) for _ = 1:N], [randn(K, 3natoms) for _ = 1:N],
                                                  zeros(K), zeros(1), [1.0], [1.0], Symmetric(zeros(K, K)))
    for (int, λ) in ((false, 0.0), (true, 0.01))
        learn!(lp, [100.0, 1.0], int; λ = λ)
        Bt, dBt = reduce(hcat, lp.B)', reduce(hcat, lp.dB)'
        A = int ? hcat([ones(N); zeros(size(dBt, 1))], [Bt; dBt]) : [Bt; dBt]
        b = [lp.e; reduce(vcat, lp.f)]
        Q = Diagonal([100.0 * ones(N); ones(size(dBt, 1))])
        βs = (A' * Q * A + λ * I) \ (A' * Q * b)
        got = int ? [lp.β0; lp.β] : lp.β
        @test relerr(got, βs) < 1e-9
    end
    M, v = B.normal_matrices(lp, [100.0, 1.0])
    A, W, b = PotentialLearning.assemble_matrices(lp, [100.0, 1.0])                      # linear-learn.jl:286-299
    @test relerr(M, A' * W * A) < 1e-10 && relerr(v, A' * W * b) < 1e-10

    # compute_eigen / PCAState (dimension_reduction.jl:11-14, pca_state.jl:20-40)
    rows = [randn(12) .+ 10.0 for _ = 1:500]
    λ, ϕ = PotentialLearning.compute_eigen(rows)
    Qref = Symmetric(mean(di * di' for di in rows))
    @test relerr(λ, eigen(Qref).values) < 1e-10
    pca = PCAState(tol = 3)
    ld2 = [LocalDescriptors([randn(12) .+ 3.0 for _ = 1:num_atoms]) for _ = 1:200]
    ds2 = DataSet(Configuration.(ld2))
    fit!(ds2, pca)
    dd = sum.(get_values.(get_local_descriptors.(ds2)))
    m = sum(dd) / length(dd)
    @test relerr(pca.m, m) < 1e-12
    λr, _ = eigen(Symmetric(mean((x .- m) * (x .- m)' for x in dd)))
    @test relerr(pca.λ, λr[end:-1:1]) < 1e-10 && size(pca.W) == (12, 3)
end
