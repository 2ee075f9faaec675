# PotentialLearningB200.jl — the Julia side of the drop-in: `using PotentialLearning, PotentialLearningB200` re-points the hot
# path of PotentialLearning.jl (v0.2.7) at libplb200.so (hand-written sm_100a CUDA, C ABI in include/plb200.h).
#
# It ONLY adds more specific methods for the signatures below; everything else of PotentialLearning stays as is:
#   KernelMatrix(F::Vector{Vector{Float64}}, k::DotProduct)                    src/Kernels/kernels.jl:211-223
#   KernelMatrix(F::Vector{Vector{Float64}}, k::RBF)   (k.d isa Euclidean)     src/Kernels/kernels.jl:211-223
#   KernelMatrix(F::Vector{Vector{Float64}}, k::InverseMultiquadric)           src/Kernels/kernels.jl:139-164, 211-223
#   KernelMatrix(F1, F2, k)                                                    src/Kernels/kernels.jl:229-244
#   learn!(lp::UnivariateLinearProblem, ws::Vector, int::Bool)                 src/Learning/linear-learn.jl:178-215
#   learn!(lp::CovariateLinearProblem,  ws::Vector, int::Bool; λ)              src/Learning/linear-learn.jl:226-268
#   learn!(lp::UnivariateLinearProblem, α::Real)   (assembly only)             src/Learning/linear-learn.jl:11-23
#   compute_eigen(d::Vector{<:Vector{<:Real}})                                 src/DimensionReduction/dimension_reduction.jl:11-14
#   fit!(ds::DataSet, pca::PCAState)                                           src/DimensionReduction/pca_state.jl:20-40
#   compute_features(ds::DataSet, f::GlobalMean / ::GlobalSum / ::CorrelationMatrix)   src/Kernels/features.jl:19-50, 76-78   (widened rows f2, f4)
#   compute_divergence(x, ::KernelSteinDiscrepancy)   (RBF over Euclidean)    src/Kernels/divergences.jl:27-49       (widened row f4)
# kDPP(ds, f, k) / kDPP(features, k) (src/SubsetSelection/dpp.jl:18-44) need no override: they call KernelMatrix and hand
# the Symmetric result to Determinantal.jl exactly as before. Optional (widened row f1): `ell_ensemble_gpu(F, k)` returns the
# eigendecomposition computed on the device for callers that construct Determinantal's ensemble from (λ, U) themselves.
#
# NOTE: Julia is not installed in the build image nor on the benchmark box, so this file has never been executed there.
# It is deliberately declarative (pack -> ccall -> wrap); all logic lives behind the C ABI, which the Python ctypes mirror
# (plb200/_lib.py) exercises symbol for symbol in tests/. There is no CPU fallback: a nonzero status raises.
module PotentialLearningB200

using LinearAlgebra
using PotentialLearning
import PotentialLearning: KernelMatrix, learn!, fit!, compute_eigen, compute_features, DotProduct, RBF, Euclidean, GlobalMean, GlobalSum,
                          UnivariateLinearProblem, CovariateLinearProblem, PCAState, DataSet, get_values, get_local_descriptors,
                          select_eigendirections

const libplb200 = get(ENV, "PLB200_LIB", joinpath(@__DIR__, "..", "plb200", "libplb200.so"))

const PL_KERNEL_DOT, PL_KERNEL_RBF = Cint(1), Cint(2)
const PL_CINV_IDENTITY, PL_CINV_DIAGONAL, PL_CINV_FULL = Cint(0), Cint(1), Cint(2)
const PL_FILL_JULIA_U = Cint(2)   # Symmetric(K) with uplo = :U over column-major storage

function check(rc::Cint, what)
    rc == 0 || error("$what failed (code $rc): ", unsafe_string(ccall((:pl_last_error, libplb200), Cstring, ())))
    nothing
end

function __init__()
    dev = parse(Cint, get(ENV, "PLB200_DEVICE", "-1"))
    check(ccall((:pl_init, libplb200), Cint, (Cint,), dev), "pl_init")   # errors without an sm_100 GPU: no CPU path
    atexit(() -> ccall((:pl_finalize, libplb200), Cvoid, ()))
end

# zeros(n, n) in pinned host memory owned by the library: D2H into it runs at the PCIe rate (config 2: 36 ms end to end instead of 105 ms
# into pageable memory, 510 ms into freshly calloc'ed pages). Set PLB200_PINNED_RESULTS=1 to make KernelMatrix allocate its result this way.
function pinned_zeros(n::Integer, m::Integer)
    p = ccall((:pl_host_alloc, libplb200), Ptr{Cvoid}, (Int64,), 8 * n * m)
    p == C_NULL && error("pl_host_alloc failed: ", unsafe_string(ccall((:pl_last_error, libplb200), Cstring, ())))
    A = unsafe_wrap(Array, Ptr{Float64}(p), (n, m))
    finalizer(_ -> ccall((:pl_host_free, libplb200), Cint, (Ptr{Cvoid},), p), A)
    A
end
result_zeros(n, m) = get(ENV, "PLB200_PINNED_RESULTS", "0") == "1" ? pinned_zeros(n, m) : zeros(n, m)

# Vector{Vector{Float64}} -> d x n column-major == the library's row-major n x d (zero transposition)
pack(F::Vector{Vector{Float64}}) = reduce(hcat, F)

cinv_args(e::Euclidean) =
    e.Cinv isa Diagonal ? (all(==(1.0), e.Cinv.diag) ? (PL_CINV_IDENTITY, C_NULL, nothing) : (PL_CINV_DIAGONAL, pointer(e.Cinv.diag), e.Cinv.diag)) :
                          (let M = Matrix(e.Cinv); (PL_CINV_FULL, pointer(M), M) end)

# ---- KernelMatrix(F, k) -> Symmetric{Float64,Matrix{Float64}} (kernels.jl:211-223) -------------------------------------------
function KernelMatrix(F::Vector{Vector{Float64}}, k::DotProduct)
    X = pack(F); d, n = size(X)
    K = result_zeros(n, n)                                                             # kernels.jl:216
    check(ccall((:pl_gram_dot, libplb200), Cint, (Ptr{Cdouble}, Int64, Int64, Int64, Cint, Ptr{Cdouble}, Int64, Cint),
                X, n, d, d, Cint(k.α), K, n, PL_FILL_JULIA_U), "pl_gram_dot")
    Symmetric(K)                                                                       # kernels.jl:222
end

function KernelMatrix(F::Vector{Vector{Float64}}, k::RBF)
    k.d isa Euclidean || return invoke(KernelMatrix, Tuple{Vector{Vector{Float64}},PotentialLearning.Kernel}, F, k)
    X = pack(F); d, n = size(X)
    kind, cptr, keep = cinv_args(k.d)
    K = result_zeros(n, n)
    GC.@preserve keep check(ccall((:pl_gram_rbf, libplb200), Cint,
                (Ptr{Cdouble}, Int64, Int64, Int64, Cint, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cdouble}, Int64, Cint),
                X, n, d, d, kind, cptr, Float64(k.ℓ), Float64(k.β), K, n, PL_FILL_JULIA_U), "pl_gram_rbf")
    Symmetric(K)
end

function KernelMatrix(F::Vector{Vector{Float64}}, k::PotentialLearning.InverseMultiquadric)     # kernels.jl:139-164 (widened row)
    k.d isa Euclidean || return invoke(KernelMatrix, Tuple{Vector{Vector{Float64}},PotentialLearning.Kernel}, F, k)
    X = pack(F); d, n = size(X)
    kind, cptr, keep = cinv_args(k.d)
    K = result_zeros(n, n)
    GC.@preserve keep check(ccall((:pl_gram_imq, libplb200), Cint,
                (Ptr{Cdouble}, Int64, Int64, Int64, Cint, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cdouble}, Int64, Cint),
                X, n, d, d, kind, cptr, Float64(k.c2), Float64(k.ℓ), K, n, PL_FILL_JULIA_U), "pl_gram_imq")
    Symmetric(K)
end

# ---- KernelMatrix(F1, F2, k) -> Matrix m x n (kernels.jl:229-244). Column-major m x n == row-major n x m: pass F2 first. ------
function KernelMatrix(F1::Vector{Vector{Float64}}, F2::Vector{Vector{Float64}}, k::Union{DotProduct,RBF})
    (k isa RBF && !(k.d isa Euclidean)) && return invoke(KernelMatrix, Tuple{Vector{Vector{Float64}},Vector{Vector{Float64}},PotentialLearning.Kernel}, F1, F2, k)
    X1 = pack(F1); X2 = pack(F2); d, m = size(X1); n = size(X2, 2)
    K = zeros(m, n)
    kid, α, kind, cptr, keep, ℓ, β = k isa DotProduct ? (PL_KERNEL_DOT, Cint(k.α), PL_CINV_IDENTITY, C_NULL, nothing, 1.0, 1.0) :
                                                      (PL_KERNEL_RBF, Cint(0), cinv_args(k.d)..., Float64(k.ℓ), Float64(k.β))
    GC.@preserve keep check(ccall((:pl_gram_cross, libplb200), Cint,
                (Cint, Ptr{Cdouble}, Int64, Int64, Ptr{Cdouble}, Int64, Int64, Int64, Cint, Cint, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cdouble}, Int64),
                kid, X2, n, d, X1, m, d, d, α, kind, cptr, ℓ, β, K, m), "pl_gram_cross")
    K
end

# ---- normal equations ---------------------------------------------------------------------------------------------------------
# rows of A as the library wants them: [B' ; dB'] row-major R x K == hcat of the K-vectors / K x m blocks, column-major K x R
pack_rows(blocks) = reduce(hcat, blocks)

function normal_eq(A::Matrix{Float64}, b::Vector{Float64}, ne::Integer, we, wf, int::Bool, λ; w = nothing)
    K, R = size(A); n = K + (int ? 1 : 0)
    AtWA = zeros(n, n); AtWb = zeros(n)
    check(ccall((:pl_normal_eq, libplb200), Cint,
                (Ptr{Cdouble}, Int64, Int64, Int64, Ptr{Cdouble}, Int64, Cdouble, Cdouble, Ptr{Cdouble}, Cint, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}),
                A, R, K, K, w === nothing ? C_NULL : w, ne, we, wf, b, int, λ, AtWA, AtWb), "pl_normal_eq")
    AtWA, AtWb
end

solve(M, v) = try M \ v catch e; println(e); println("Linear system will be solved using pinv."); pinv(M) * v end   # linear-learn.jl:199-205

function learn!(lp::UnivariateLinearProblem, ws::Vector, int::Bool)                    # linear-learn.jl:178-215
    A = pack_rows(lp.iv_data); b = Float64.(reduce(vcat, lp.dv_data))
    M, v = normal_eq(A, b, size(A, 2), ws[1], ws[1], int, 0.0)
    βs = solve(M, v)
    if int; lp.β0 .= βs[1]; lp.β .= βs[2:end] else lp.β .= βs end
end

# Streamed form (pl_normal_eq_blocks, widened row f2): lp.dB[i] (K x 3natoms, column-major) IS 3natoms rows of K in memory, so the
# blocks are handed over as they lie — no hcat copy of the 11.5 GB design matrix; H2D chunks overlap the SYRK inside the library.
function normal_eq_blocks(blocks::Vector{<:StridedMatrix{Float64}}, b::Vector{Float64}, ne::Integer, we, wf, int::Bool, λ)
    K = size(blocks[1], 1); n = K + (int ? 1 : 0)
    ptrs = [pointer(B) for B in blocks]; rows = Int64[size(B, 2) for B in blocks]
    AtWA = zeros(n, n); AtWb = zeros(n)
    GC.@preserve blocks check(ccall((:pl_normal_eq_blocks, libplb200), Cint,
                (Ptr{Ptr{Cdouble}}, Ptr{Int64}, Int64, Int64, Ptr{Cdouble}, Int64, Cdouble, Cdouble, Ptr{Cdouble}, Cint, Cdouble, Int64, Ptr{Cdouble}, Ptr{Cdouble}),
                ptrs, rows, length(blocks), K, C_NULL, ne, we, wf, b, int, λ, 0, AtWA, AtWb), "pl_normal_eq_blocks")
    AtWA, AtWb
end

function learn!(lp::CovariateLinearProblem, ws::Vector, int::Bool; λ::Real = 0.0)      # linear-learn.jl:226-268
    blocks = vcat([pack_rows(lp.B)], [Matrix{Float64}(dBi) for dBi in lp.dB])          # Matrix(...) is a no-op for Matrix{Float64}
    b = Float64.(vcat(lp.e, reduce(vcat, lp.f)))
    M, v = normal_eq_blocks(blocks, b, length(lp.e), ws[1], ws[2], int, Float64(λ))
    βs = solve(M, v)
    if int; lp.β0 .= βs[1]; lp.β .= βs[2:end] else lp.β .= βs end
end

function learn!(lp::UnivariateLinearProblem, α::Real)                                  # linear-learn.jl:11-23
    A = pack_rows(lp.iv_data); b = Float64.(reduce(vcat, lp.dv_data))
    AtA, Atb = normal_eq(A, b, 0, 1.0, 1.0, false, 0.0)                                # sum(v*v'), sum(v*b) on the GPU (:16-17)
    Q = pinv(AtA, α)
    lp.β .= Q * Atb
    lp.σ .= std(Atb - AtA * lp.β)
    lp.Σ .= Symmetric(lp.σ[1]^2 * Q)
end

# ---- covariance ---------------------------------------------------------------------------------------------------------------
function cov_gpu(d::Vector{<:Vector{<:Real}}; center = false, mean = Float64[])
    X = reduce(hcat, d); K, n = size(X)
    C = zeros(K, K); m = isempty(mean) ? zeros(K) : Float64.(mean)
    check(ccall((:pl_cov, libplb200), Cint, (Ptr{Cdouble}, Int64, Int64, Int64, Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}),
                Float64.(X), n, K, K, m, !isempty(mean), center, C), "pl_cov")
    Symmetric(C), m
end

# eigen(Symmetric(Q)) on the device (pl_eigh, widened row f1): row k of the library's V is column k of Julia's `vectors`
function eigen_gpu(Q::Symmetric{Float64,Matrix{Float64}})
    n = size(Q, 1); W = zeros(n); V = zeros(n, n)
    A = Q.data                                                                          # uplo = :U of column-major == PL_FILL_RM_LOWER
    check(ccall((:pl_eigh, libplb200), Cint, (Ptr{Cdouble}, Int64, Int64, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Int64),
                A, n, n, Q.uplo == 'U' ? Cint(2) : Cint(1), W, V, n), "pl_eigh")
    Eigen(W, V)
end

compute_eigen(d::Vector{T}) where {T<:Vector{<:Real}} = eigen_gpu(first(cov_gpu(d)))    # dimension_reduction.jl:11-14

# EllEnsemble(KernelMatrix(F, k)) fused on the device (pl_ell_ensemble): (λ ascending, K as Symmetric). U stays resident in the
# library; `ell_inclusion_prob(α, n)` and `ell_eigvecs(idx, n)` read it there (Determinantal.inclusion_prob / the eigenvectors a
# k-DPP draw selected), so n x n never crosses PCIe.
function ell_ensemble_gpu(F::Vector{Vector{Float64}}, k::Union{DotProduct,RBF})
    X = pack(F); d, n = size(X); λ = zeros(n); K = zeros(n, n)
    kid, α, kind, cptr, keep, ℓ, β = k isa DotProduct ? (PL_KERNEL_DOT, Cint(k.α), PL_CINV_IDENTITY, C_NULL, nothing, 1.0, 1.0) :
                                                      (PL_KERNEL_RBF, Cint(0), cinv_args(k.d)..., Float64(k.ℓ), Float64(k.β))
    GC.@preserve keep check(ccall((:pl_ell_ensemble, libplb200), Cint,
                (Cint, Ptr{Cdouble}, Int64, Int64, Int64, Cint, Cint, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Int64),
                kid, X, n, d, d, α, kind, cptr, ℓ, β, λ, K, n), "pl_ell_ensemble")
    λ, Symmetric(K)
end
ell_inclusion_prob(α::Real, n::Integer) = (p = zeros(n);
    check(ccall((:pl_ell_inclusion_prob, libplb200), Cint, (Cdouble, Int64, Ptr{Cdouble}), Float64(α), n, p), "pl_ell_inclusion_prob"); p)
ell_eigvecs(idx::Vector{Int64}, n::Integer) = (V = zeros(n, length(idx));                # 0-based eigenvector indices; column j = U[:, idx[j]+1]
    check(ccall((:pl_ell_eigvecs, libplb200), Cint, (Ptr{Int64}, Int64, Int64, Ptr{Cdouble}, Int64), idx, length(idx), n, V, n), "pl_ell_eigvecs"); V)

# device scratch held by the library / hand it back (keep_resident = false also drops the resident L-ensemble and an open accumulator)
workspace_bytes() = ccall((:pl_workspace_bytes, libplb200), Int64, ())
workspace_trim(keep_resident::Bool = true) = check(ccall((:pl_workspace_trim, libplb200), Cint, (Cint,), keep_resident ? 1 : 0), "pl_workspace_trim")

# first phase of a k-DPP draw (which eigenvectors; Kulesza & Taskar Alg. 8) on the device: `αλ` the rescaled eigenvalues,
# `esp_select(αλ, k, rand(rng, length(αλ)))` -> 0-based eigenvector indices for `ell_sample`
function esp_select(αλ::Vector{Float64}, k::Integer, uniforms::Vector{Float64})
    idx = zeros(Int64, k); count = Ref{Int64}(0)
    check(ccall((:pl_esp_select, libplb200), Cint, (Ptr{Cdouble}, Int64, Int64, Ptr{Cdouble}, Ptr{Int64}, Ptr{Int64}), αλ, length(αλ), k, uniforms, idx, count), "pl_esp_select")
    idx[1:count[]]
end

# second phase of a k-DPP draw on the resident ensemble: projection-DPP sampling over the eigenvectors `idx` (0-based) that the first phase
# (Determinantal's ESP selection) picked; the uniforms come from the caller's RNG: `ell_sample(idx, n, rand(rng, length(idx)))` -> 1-based items
function ell_sample(idx::Vector{Int64}, n::Integer, uniforms::Vector{Float64})
    items = zeros(Int64, length(idx))
    check(ccall((:pl_ell_sample, libplb200), Cint, (Ptr{Int64}, Int64, Int64, Ptr{Cdouble}, Ptr{Int64}), idx, length(idx), n, uniforms, items), "pl_ell_sample")
    items .+ 1
end

# greedy MAP of the resident ensemble (Determinantal.greedy_subset, dpp.jl:60): 1-based items, shorter than k if the rank is exhausted
function ell_greedy(k::Integer, n::Integer)
    items = zeros(Int64, k)
    check(ccall((:pl_ell_greedy, libplb200), Cint, (Int64, Int64, Ptr{Int64}), k, n, items), "pl_ell_greedy")
    [i + 1 for i in items if i >= 0]
end

# ---- GlobalMean / GlobalSum features of a whole DataSet (features.jl:19-34, 76-78) in one segmented reduction ------------------
function global_features(ds::DataSet, mean::Bool)
    locals = [reduce(hcat, get_values(get_local_descriptors(c))) for c in ds]            # d x natoms_i each == natoms_i rows of d
    L = reduce(hcat, locals); d = size(L, 1); N = length(locals)
    offsets = Int64[0; cumsum(size.(locals, 2))]
    X = zeros(d, N)
    check(ccall((:pl_global_feature, libplb200), Cint, (Ptr{Cdouble}, Int64, Ptr{Int64}, Int64, Int64, Cint, Ptr{Cdouble}, Int64),
                L, d, offsets, N, d, mean, X, d), "pl_global_feature")
    [X[:, i] for i in 1:N]                                                              # Vector{Vector{Float64}} like compute_feature.(ds, ...)
end
compute_features(ds::DataSet, f::GlobalMean; dt = nothing) = global_features(ds, true)
compute_features(ds::DataSet, f::GlobalSum; dt = nothing) = global_features(ds, false)

function fit!(ds::DataSet, pca::PCAState)                                              # pca_state.jl:20-40 (as written: descriptors only)
    d = sum.(get_values.(get_local_descriptors.(ds)))
    Q, m = cov_gpu(d; center = true, mean = pca.m == [] ? Float64[] : pca.m)            # mean persists (:34)
    pca.m = m
    λ, ϕ = eigen_gpu(Q); λ, ϕ = λ[end:-1:1], ϕ[:, end:-1:1]                             # dimension_reduction.jl:16-17
    Σ = 1.0 .- cumsum(λ) / sum(λ)
    pca.λ, pca.W = λ, (pca.tol isa Int ? ϕ[:, 1:pca.tol] : ϕ[:, Σ .> pca.tol])
    nothing
end

compute_features(ds::DataSet, f::PotentialLearning.CorrelationMatrix; dt = nothing) = begin                  # features.jl:48-50
    locals = [reduce(hcat, get_values(get_local_descriptors(c))) for c in ds]
    L = reduce(hcat, locals); d = size(L, 1); N = length(locals)
    offsets = Int64[0; cumsum(size.(locals, 2))]
    C = zeros(d, d, N)
    check(ccall((:pl_correlation_feature, libplb200), Cint, (Ptr{Cdouble}, Int64, Ptr{Int64}, Int64, Int64, Ptr{Cdouble}),
                L, d, offsets, N, d, C), "pl_correlation_feature")
    [Symmetric(C[:, :, i]) for i in 1:N]
end

# ---- predictions of a fitted linear model: get_all_energies(ds, lb) / get_all_forces(ds, lb) (Data/utils.jl:42-65) ------------
function predict_gpu(A::Matrix{Float64}, β::Vector{Float64}, β0::Float64, ne::Integer)                       # A: K x R (rows of the design matrix as columns)
    K, R = size(A); y = zeros(R)
    check(ccall((:pl_predict, libplb200), Cint, (Ptr{Cdouble}, Int64, Int64, Int64, Ptr{Cdouble}, Cdouble, Int64, Ptr{Cdouble}),
                A, R, K, K, β, β0, ne, y), "pl_predict")
    y
end

# ---- ooc_learn! running sums (linear-learn.jl:301-391): AtWA .+= A'*W*A; AtWb .+= A'*W*b (:355-356) stay on the device ------------
# Inside ooc_learn!'s `for config in ds_train` loop: A = [global_descr'; force_descrs] (rows x K), buffer a few configurations, then
#   neq_add(permutedims(A), b, w)      # per-row weights w = [ws[1]*e_norm; ws[2]*ones(3natoms)]  (:341-352)
# and after the loop `AtWA, AtWb = neq_finish(K)`; the ridge / solve / cond print of :358-391 are unchanged.
neq_begin(K::Integer) = check(ccall((:pl_neq_begin, libplb200), Cint, (Int64,), K), "pl_neq_begin")
function neq_add(At::Matrix{Float64}, b::Vector{Float64}, w::Vector{Float64})                                   # At: K x rows
    K, R = size(At)
    check(ccall((:pl_neq_add, libplb200), Cint, (Ptr{Cdouble}, Int64, Int64, Ptr{Cdouble}, Int64, Cdouble, Cdouble, Ptr{Cdouble}),
                At, R, K, w, 0, 1.0, 1.0, b), "pl_neq_add")
end
function neq_finish(K::Integer; int::Bool = false, λ::Real = 0.0)
    n = K + (int ? 1 : 0); AtWA = zeros(n, n); AtWb = zeros(n)
    check(ccall((:pl_neq_finish, libplb200), Cint, (Cint, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}), int, Float64(λ), AtWA, AtWb), "pl_neq_finish")
    AtWA, AtWb
end

# ---- KernelSteinDiscrepancy with an RBF(Euclidean) kernel (divergences.jl:27-49) ----------------------------------------------------
function PotentialLearning.compute_divergence(x::Vector{Vector{Float64}}, div::PotentialLearning.KernelSteinDiscrepancy)
    k = div.kernel
    (k isa RBF && k.d isa Euclidean) || return invoke(PotentialLearning.compute_divergence,
                                                      Tuple{Vector{<:Union{Real,Vector{<:Real}}},PotentialLearning.KernelSteinDiscrepancy}, x, div)
    X = pack(x); S = pack(Vector{Float64}.(div.score.(x))); d, n = size(X)
    kind, cptr, keep = cinv_args(k.d)
    out = zeros(1)
    GC.@preserve keep check(ccall((:pl_ksd_rbf, libplb200), Cint,
                (Ptr{Cdouble}, Ptr{Cdouble}, Int64, Int64, Int64, Int64, Cint, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cdouble}),
                X, S, n, d, d, d, kind, cptr, Float64(k.ℓ), Float64(k.β), out), "pl_ksd_rbf")
    out[1]
end

# ---- DBSCAN's distance matrices (dbscan.jl:132-170): positions as 3 x natoms x n (column-major) == n x natoms x 3 row-major -------------
function distance_matrix_gpu(P::Array{Float64,3}; box_lengths::Union{Nothing,Vector{Float64}} = nothing)
    natoms, n = size(P, 2), size(P, 3); D = zeros(n, n)
    if box_lengths === nothing
        check(ccall((:pl_rmsd_kabsch, libplb200), Cint, (Ptr{Cdouble}, Int64, Int64, Ptr{Cdouble}, Int64), P, n, natoms, D, n), "pl_rmsd_kabsch")
    else
        check(ccall((:pl_rmsd_periodic, libplb200), Cint, (Ptr{Cdouble}, Int64, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Int64), P, n, natoms, box_lengths, D, n), "pl_rmsd_periodic")
    end
    D
end

end # module
