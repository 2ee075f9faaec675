# dimred.jl — the second moment behind compute_eigen (dimension_reduction.jl:11-14) and the mean / centring of fit!(ds, ::PCAState)
# (pca_state.jl:34-37) on the B200. `eigen` stays LAPACK in Julia by default, exactly like the reference.

"""
    covariance(d; center = false, mean = Float64[]) -> (Symmetric Q, m)

`Q = mean(di * di' for di in d)` (1/n, dimension_reduction.jl:12). `center = true` subtracts `m` first: `m = sum(d) / length(d)`
(pca_state.jl:35) when `mean` is empty, else the given (persisted) mean is used as is (pca_state.jl:34). One pass over the rows.
"""
function covariance(d::Vector{<:Vector{<:Real}}; center::Bool = false, mean = Float64[])
    X = pack(d); K = size(X, 1)
    given = !isempty(mean)
    m = given ? Vector{Float64}(mean) : zeros(K)
    C = pl_cov(X, m, given, center)
    Symmetric(C), m
end

# eigen(Symmetric(Q)) on the device — OPT-IN (PLB200_EIGEN=gpu or a direct call). Eigenvector signs / the basis inside degenerate
# clusters are the solver's, so pca.W may differ from LAPACK's by signs: that is why it is not the default.
function eigen_gpu(Q::Symmetric{Float64,Matrix{Float64}})
    W, V = pl_eigh(Q.data, Q.uplo)
    Eigen(W, V)
end
eigen_of(Q) = get(ENV, "PLB200_EIGEN", "lapack") == "gpu" ? eigen_gpu(Q) : eigen(Symmetric(Q))   # dimension_reduction.jl:13

# fit!(ds, pca::PCAState) (pca_state.jl:20-40): global-sum descriptors (force rows are never appended in the reference as written:
# the `fd` of :29 is undefined, the catch returns d), mean unless pca.m is already set, centred second moment, eigen, selection.
function fit_pcastate!(ds, pca)
    d = global_features(ds, false)                                                      # sum.(get_values.(get_local_descriptors.(ds))) (:23)
    Q, m = covariance(d; center = true, mean = pca.m == [] ? Float64[] : pca.m)
    if pca.m == []
        pca.m = m                                                                       # :34-36
    end
    λ, ϕ = eigen_of(Q)
    λ, ϕ = λ[end:-1:1], ϕ[:, end:-1:1]                                                  # dimension_reduction.jl:17, 24
    if pca.tol isa Int
        pca.λ, pca.W = λ, ϕ[:, 1:pca.tol]                                               # :26
    else
        Σ = 1.0 .- cumsum(λ) / sum(λ)
        pca.λ, pca.W = λ, ϕ[:, Σ .> pca.tol]                                            # :18-19
    end
    nothing
end
