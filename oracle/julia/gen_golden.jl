# gen_golden.jl — golden vectors from the UNMODIFIED reference (cesmix-mit/PotentialLearning.jl v0.2.7) for the hot path this
# repository replaces. TEST INFRASTRUCTURE. It loads only the reference (never PotentialLearningB200) and writes inputs-by-recipe and
# outputs-by-value that tests/test_reference_goldens.py consumes, turning the repository's "parity unpinned" into pinned parity.
#
# The build image of this repository has no Julia, so this script has not been run there: run it on any machine with Julia >= 1.10
# and the reference installed,
#
#     julia --project=/path/to/PotentialLearning.jl oracle/julia/gen_golden.jl tests/golden/julia
#
# and commit the directory it writes (about 6 MB). No GPU is needed.
#
# File format (no JLD / NPZ dependency): every array is a raw little-endian Float64 file `<name>.f64` in Julia's column-major order;
# `manifest.json` lists name -> dims for every array and the scalar parameters of every case.
# Large inputs are not stored: they come from a 64-bit LCG + sum-of-four-uniforms generator that uses only exactly rounded IEEE
# operations (integer multiply-add, one int->float conversion, additions, one multiplication by sqrt(3)), re-implemented bit for
# bit in the Python test; the manifest stores a checksum of every regenerated input.
using LinearAlgebra, Statistics, Random
using PotentialLearning

const OUT = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "..", "tests", "golden", "julia")
mkpath(OUT)

# ---- deterministic inputs ---------------------------------------------------------------------------------------------------------
mutable struct LCG
    s::UInt64
end
next!(g::LCG) = (g.s = g.s * 0x5851f42d4c957f2d + 0x14057b7ef767814f; g.s)            # Knuth MMIX constants, arithmetic mod 2^64
unif!(g::LCG) = Float64(next!(g) >> 11) * 2.0^-53                                      # exact: 53 random bits
gauss!(g::LCG) = (((unif!(g) + unif!(g)) + unif!(g)) + unif!(g) - 2.0) * sqrt(3.0)     # mean 0, variance 1 (sum of four uniforms)
# d x n matrix filled column by column (== the library's row-major n x d): X[k, i] = shift[k] + scale * gauss
function fill_matrix(seed::Integer, d::Int, n::Int, shift::Vector{Float64}, scale::Float64)
    g = LCG(UInt64(seed))
    X = Matrix{Float64}(undef, d, n)
    for i = 1:n, k = 1:d
        X[k, i] = shift[k] + scale * gauss!(g)
    end
    X
end
fill_vector(seed::Integer, n::Int, scale::Float64) = vec(fill_matrix(seed, 1, n, [0.0], scale))
probe(X) = [X[1], X[(length(X) + 1) ÷ 2], X[end]]      # exact spot values of a regenerated input (first, middle, last in memory order)

# ---- output ---------------------------------------------------------------------------------------------------------------------------
const ARRAYS = Dict{String,Any}()
const CASES = Vector{String}()
function put(name::String, A)
    B = Float64.(A isa Number ? [A] : collect(A))
    open(joinpath(OUT, name * ".f64"), "w") do io
        write(io, htol.(vec(B)))
    end
    ARRAYS[name] = collect(size(B))
    nothing
end
jstr(x::AbstractString) = "\"" * x * "\""
jval(x::AbstractString) = jstr(x)
jval(x::Bool) = x ? "true" : "false"
jval(x::Integer) = string(x)
jval(x::AbstractFloat) = repr(Float64(x))
jval(x::AbstractVector) = "[" * join(jval.(x), ", ") * "]"
function case(name::String; kw...)
    push!(CASES, "  " * jstr(name) * ": {" * join([jstr(string(k)) * ": " * jval(v) for (k, v) in kw], ", ") * "}")
end
sampled_rows(n, count, seed) = sort(unique([Int(next!(LCG(UInt64(seed + 977 * t))) % UInt64(n)) + 1 for t = 1:count]))   # 1-based

# =====================================================================================================================================
# 1. test/kernels/kernel_tests.jl fixture (10 configs x 20 atoms x d = 8): GlobalMean / GlobalSum features, both KernelMatrix forms,
#    DotProduct and RBF(Euclidean) with identity, Diagonal and Symmetric Cinv  (features.jl:19-34, kernels.jl:30-36, 68-75, 211-244)
# =====================================================================================================================================
let d = 8, natoms = 20, ncfg = 10
    L = fill_matrix(101, d, natoms * ncfg, 0.5 .* collect(1.0:d), 1.0)
    ld = [LocalDescriptors([L[:, (c - 1) * natoms + a] for a = 1:natoms]) for c = 1:ncfg]
    ds = DataSet(Configuration.(ld))
    f_gm = compute_feature.(ld, (GlobalMean(),))
    f_gs = compute_feature.(ld, (GlobalSum(),))
    put("kt_locals", L); put("kt_global_mean", reduce(hcat, f_gm)); put("kt_global_sum", reduce(hcat, f_gs))
    put("kt_K_dot", Matrix(KernelMatrix(f_gm, DotProduct())))                        # full symmetric values
    put("kt_K_dot_storage", KernelMatrix(f_gm, DotProduct()).data)                   # zeros(n,n) + :U triangle as stored (kernels.jl:216-222)
    put("kt_K_dot3", Matrix(KernelMatrix(f_gm, DotProduct(3))))
    put("kt_K_rbf", Matrix(KernelMatrix(f_gm, RBF(Euclidean(d)))))
    put("kt_K_rbf_ds", Matrix(KernelMatrix(ds, GlobalMean(), RBF(Euclidean(d); ℓ = 1.7, β = 0.6))))
    cdiag = [0.5 + 0.25 * k for k = 1:d]
    put("kt_cinv_diag", cdiag)
    put("kt_K_rbf_diag", Matrix(KernelMatrix(f_gm, RBF(Euclidean(Diagonal(cdiag))))))
    R = fill_matrix(102, d, d, zeros(d), 1.0); C = Symmetric(R * R' / d + I)
    put("kt_cinv_full", Matrix(C))
    put("kt_K_rbf_full", Matrix(KernelMatrix(f_gm, RBF(Euclidean(C)))))
    put("kt_K_cross_rbf", KernelMatrix(f_gm[1:4], f_gm[5:10], RBF(Euclidean(d))))     # 4 x 6 (kernels.jl:229-244)
    put("kt_K_cross_dot", KernelMatrix(f_gm[1:4], f_gm[5:10], DotProduct()))
    case("kernel_tests"; d = d, natoms = natoms, ncfg = ncfg, seed = 101, ell = 1.7, beta = 0.6)
end

# =====================================================================================================================================
# 2. BASELINE configs[0]: kDPP kernel build, N = 1000 configurations x 32 atoms x d = 26, GlobalMean, DotProduct(alpha = 2)
# =====================================================================================================================================
let d = 26, natoms = 32, N = 1000
    mu = 2.0 .* fill_vector(201, d, 1.0)
    L = fill_matrix(202, d, natoms * N, mu, 1.0)
    F = [mean([L[:, (c - 1) * natoms + a] for a = 1:natoms]) for c = 1:N]            # compute_feature(::GlobalMean) = mean(get_values(B)) (features.jl:33)
    K = KernelMatrix(F, DotProduct())
    rows = sampled_rows(N, 32, 203)
    put("c1_mu", mu); put("c1_features", reduce(hcat, F)); put("c1_rows", rows); put("c1_K_rows", Matrix(K)[rows, :])
    put("c1_K_trace_sum", [tr(K), sum(Matrix(K))])
    case("config1"; d = d, natoms = natoms, N = N, seed_mu = 201, seed_locals = 202, input_checksum = sum(L), input_probe = probe(L))
end

# =====================================================================================================================================
# 3. BASELINE configs[1] sub-sampled: RBF(Euclidean(300)), N = 3000 (of 20 000), features mu + N(0, 1/96)
# =====================================================================================================================================
let d = 300, N = 3000
    mu = 2.0 .* fill_vector(301, d, 1.0)
    X = fill_matrix(302, d, N, mu, 1.0 / sqrt(96.0))
    F = [X[:, i] for i = 1:N]
    K = KernelMatrix(F, RBF(Euclidean(d)))
    rows = sampled_rows(N, 16, 303)
    put("c2_mu", mu); put("c2_rows", rows); put("c2_K_rows", Matrix(K)[rows, :]); put("c2_K_trace_sum", [tr(K), sum(Matrix(K))])
    case("config2_sub"; d = d, N = N, seed_mu = 301, seed_x = 302, input_checksum = sum(X), input_probe = probe(X), ell = 1.0, beta = 1.0)
end

# =====================================================================================================================================
# 4. learn!: test/learning/linear_tests.jl shapes and BASELINE configs[2] sub-sampled (200 of 10 000 configurations x 96 atoms, K = 500)
#    WLS (linear-learn.jl:226-268) with intercept and ridge, the products A'QA + lambda I, A'Qb themselves, OLS sums (:16-17),
#    per-row weights in the ooc_learn! form (:341-356)
# =====================================================================================================================================
function covariate_problem(seed, N, natoms, K)
    Bm = fill_matrix(seed, K, N, zeros(K), sqrt(Float64(natoms)))
    dBm = fill_matrix(seed + 1, K, 3 * natoms * N, zeros(K), 1.0)
    e = fill_vector(seed + 2, N, 1.0)
    f = fill_vector(seed + 3, 3 * natoms * N, 1.0)
    lp = PotentialLearning.CovariateLinearProblem(e, [f[(c - 1) * 3natoms + 1:c * 3natoms] for c = 1:N], [Bm[:, c] for c = 1:N],
                                                  [dBm[:, (c - 1) * 3natoms + 1:c * 3natoms] for c = 1:N], zeros(K), zeros(1), [1.0], [1.0],
                                                  Symmetric(zeros(K, K)))
    lp, sum(Bm) + sum(dBm) + sum(e) + sum(f)
end
for (tag, N, natoms, K, seed) in (("lt", 100, 20, 8, 401), ("c3", 200, 96, 500, 411))
    lp, chk = covariate_problem(seed, N, natoms, K)
    ws = [100.0, 1.0]
    for (int, λ, sfx) in ((false, 0.0, "plain"), (true, 0.01, "int_ridge"))
        learn!(lp, ws, int; λ = λ)                                                       # the reference's own method
        put("$(tag)_beta_$(sfx)", lp.β); put("$(tag)_beta0_$(sfx)", lp.β0)
        A, W, b = PotentialLearning.assemble_matrices(lp, ws)                             # linear-learn.jl:286-299
        if int
            A = hcat([ones(N); zeros(size(A, 1) - N)], A)                                 # :239-241
        end
        put("$(tag)_AtWA_$(sfx)", A' * W * A + λ * I); put("$(tag)_AtWb_$(sfx)", A' * W * b)   # :253
    end
    case(tag == "lt" ? "linear_tests" : "config3_sub"; N = N, natoms = natoms, K = K, seed = seed, we = 100.0, wf = 1.0, lambda = 0.01, input_checksum = chk)
end
let N = 50, natoms = 7, K = 12
    lp, chk = covariate_problem(421, N, natoms, K)
    AtAe = sum(b * b' for b in lp.B); Atbe = sum(b * e for (b, e) in zip(lp.B, lp.e))             # :45-46 (== :16-17)
    AtAf = sum(db * db' for db in lp.dB); Atbf = sum(db * f for (db, f) in zip(lp.dB, lp.f))      # :48-49
    put("ols_AtAe", AtAe); put("ols_Atbe", Atbe); put("ols_AtAf", AtAf); put("ols_Atbf", Atbf)
    # ooc_learn! accumulation with eweight_normalized = :squared (:341-356): per configuration A = [B'; dB'], W = Diagonal([ws1/natoms^2; ws2 * 1])
    ws = [30.0, 1.0]
    AtWA = zeros(K, K); AtWb = zeros(K)
    for c = 1:N
        A = [lp.B[c]'; lp.dB[c]']; b = [lp.e[c]; lp.f[c]]
        W = Diagonal([ws[1] / natoms^2; ws[2] * ones(3natoms)])
        AtWA .+= A' * W * A; AtWb .+= A' * W * b
    end
    put("ooc_AtWA", AtWA); put("ooc_AtWb", AtWb)
    case("ols_ooc"; N = N, natoms = natoms, K = K, seed = 421, we = 30.0, wf = 1.0, input_checksum = chk)
end

# =====================================================================================================================================
# 5. compute_eigen / PCAState: dimension_reduction.jl:11-28, pca_state.jl:34-38 — small (full Q, eigenvalues, |W|) and BASELINE
#    configs[3] sub-sampled (n = 20 000 of 10^7 rows, K = 1000, mean/sigma = 10)
# =====================================================================================================================================
for (tag, n, K, seed) in (("pca_small", 500, 12, 501), ("c4", 20_000, 1000, 511))
    shift = 10.0 .* fill_vector(seed, K, 1.0)
    X = fill_matrix(seed + 1, K, n, shift, 1.0)
    d = [X[:, i] for i = 1:n]
    λu, _ = PotentialLearning.compute_eigen(d)                                            # un-centred (PCA / ActiveSubspace use it so)
    m = sum(d) / length(d)                                                                # pca_state.jl:35
    dm = d .- [m]                                                                         # :37
    Q = Symmetric(mean(di * di' for di in dm))                                            # dimension_reduction.jl:12
    λ, W = PotentialLearning.select_eigendirections(dm, 3)                                # :22-28
    put("$(tag)_mean", m); put("$(tag)_lambda_uncentred", λu); put("$(tag)_lambda", λ); put("$(tag)_absW", abs.(W))
    if K <= 64
        put("$(tag)_Q", Matrix(Q))
    else
        idx = sampled_rows(K, 64, seed + 2)
        put("$(tag)_Q_idx", idx); put("$(tag)_Q_block", Matrix(Q)[idx, idx]); put("$(tag)_Q_trace", tr(Q))
    end
    case(tag == "c4" ? "config4_sub" : "pca_small"; n = n, K = K, seed = seed, input_checksum = sum(X), input_probe = probe(X))
end

open(joinpath(OUT, "manifest.json"), "w") do io
    println(io, "{")
    println(io, " \"reference\": \"PotentialLearning.jl $(pkgversion(PotentialLearning))\", \"julia\": \"$(VERSION)\", \"blas_threads\": $(BLAS.get_num_threads()),")
    println(io, " \"cases\": {\n" * join(CASES, ",\n") * "\n },")
    println(io, " \"arrays\": {\n" * join(["  " * jstr(k) * ": " * jval(v) for (k, v) in sort(collect(ARRAYS); by = first)], ",\n") * "\n }")
    println(io, "}")
end
println("wrote ", length(ARRAYS), " arrays and ", length(CASES), " cases to ", OUT)
