# lib.jl — the C ABI of libplb200.so (include/plb200.h), one thin function per entry point. All `ccall`s of the package live here.
# Conventions (include/plb200.h): descriptor matrices are ROW-major n x d == Julia's column-major d x n (`reduce(hcat, F)`), so nothing
# is transposed on either side; every function returns 0 or an error code whose message is pl_last_error().

const libplb200 = get(ENV, "PLB200_LIB", normpath(joinpath(@__DIR__, "..", "..", "..", "plb200", "libplb200.so")))

const PL_KERNEL_LINEAR, PL_KERNEL_DOT, PL_KERNEL_RBF, PL_KERNEL_IMQ = Cint(0), Cint(1), Cint(2), Cint(3)
const PL_CINV_IDENTITY, PL_CINV_DIAGONAL, PL_CINV_FULL = Cint(0), Cint(1), Cint(2)
const PL_FILL_JULIA_U = Cint(2)   # Symmetric(K) with uplo = :U over column-major storage (kernels.jl:222)

last_error() = unsafe_string(ccall((:pl_last_error, libplb200), Cstring, ()))
function check(rc::Cint, what)
    rc == 0 || error("$what failed (code $rc): ", last_error())
    nothing
end

# ---- lifetime ---------------------------------------------------------------------------------------------------------------------
pl_init(dev::Cint) = check(ccall((:pl_init, libplb200), Cint, (Cint,), dev), "pl_init")
pl_init_all(ids::Vector{Cint}) = check(ccall((:pl_init_all, libplb200), Cint, (Cint, Ptr{Cint}), Cint(length(ids)), isempty(ids) ? C_NULL : ids), "pl_init_all")
pl_finalize() = ccall((:pl_finalize, libplb200), Cvoid, ())
device_count() = Int(ccall((:pl_device_count, libplb200), Cint, ()))
workspace_bytes() = ccall((:pl_workspace_bytes, libplb200), Int64, ())
workspace_trim(keep_resident::Bool = true) = check(ccall((:pl_workspace_trim, libplb200), Cint, (Cint,), keep_resident ? 1 : 0), "pl_workspace_trim")

function last_timing()
    k, h, d = Ref(0.0), Ref(0.0), Ref(0.0)
    ccall((:pl_last_timing, libplb200), Cint, (Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), k, h, d)
    (kernel_ms = k[], h2d_ms = h[], d2h_ms = d[])
end

# zeros(n, m) in pinned host memory owned by the library: D2H into it runs at the PCIe rate. Results default to ordinary zeros(n, m)
# (the library stages pageable memory through its own pinned ring); PLB200_PINNED_RESULTS=1 makes KernelMatrix allocate its result here.
function pinned_zeros(n::Integer, m::Integer)
    p = ccall((:pl_host_alloc, libplb200), Ptr{Cvoid}, (Int64,), 8 * n * m)
    p == C_NULL && error("pl_host_alloc failed: ", last_error())
    A = unsafe_wrap(Array, Ptr{Float64}(p), (n, m))
    finalizer(_ -> ccall((:pl_host_free, libplb200), Cint, (Ptr{Cvoid},), p), A)
    A
end
result_zeros(n, m) = get(ENV, "PLB200_PINNED_RESULTS", "0") == "1" ? pinned_zeros(n, m) : zeros(n, m)

# ---- kernel matrices (kernels.jl:211-244) -----------------------------------------------------------------------------------------
# X: d x n column-major (one feature per column). K: n x n, only Julia's upper triangle is written (zeros(n,n) elsewhere, kernels.jl:216).
pl_gram_dot!(K::Matrix{Float64}, X::Matrix{Float64}, α::Integer) =
    check(ccall((:pl_gram_dot, libplb200), Cint, (Ptr{Cdouble}, Int64, Int64, Int64, Cint, Ptr{Cdouble}, Int64, Cint),
                X, size(X, 2), size(X, 1), size(X, 1), Cint(α), K, size(K, 1), PL_FILL_JULIA_U), "pl_gram_dot")

pl_gram_rbf!(K::Matrix{Float64}, X::Matrix{Float64}, kind::Cint, cinv, ℓ::Float64, β::Float64) =
    check(ccall((:pl_gram_rbf, libplb200), Cint, (Ptr{Cdouble}, Int64, Int64, Int64, Cint, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cdouble}, Int64, Cint),
                X, size(X, 2), size(X, 1), size(X, 1), kind, cinv, ℓ, β, K, size(K, 1), PL_FILL_JULIA_U), "pl_gram_rbf")

pl_gram_imq!(K::Matrix{Float64}, X::Matrix{Float64}, kind::Cint, cinv, c2::Float64, ℓ::Float64) =
    check(ccall((:pl_gram_imq, libplb200), Cint, (Ptr{Cdouble}, Int64, Int64, Int64, Cint, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cdouble}, Int64, Cint),
                X, size(X, 2), size(X, 1), size(X, 1), kind, cinv, c2, ℓ, K, size(K, 1), PL_FILL_JULIA_U), "pl_gram_imq")

# K: m x n column-major == the library's row-major n x m: F2 goes in as the library's X1, F1 as X2
pl_gram_cross!(K::Matrix{Float64}, X1::Matrix{Float64}, X2::Matrix{Float64}, kid::Cint, α::Cint, kind::Cint, cinv, ℓ::Float64, β::Float64) =
    check(ccall((:pl_gram_cross, libplb200), Cint,
                (Cint, Ptr{Cdouble}, Int64, Int64, Ptr{Cdouble}, Int64, Int64, Int64, Cint, Cint, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cdouble}, Int64),
                kid, X2, size(X2, 2), size(X2, 1), X1, size(X1, 2), size(X1, 1), size(X1, 1), α, kind, cinv, ℓ, β, K, size(K, 1)), "pl_gram_cross")

# resident, tile-packed kernel matrix (N = 200 000: 160 GB that only exists as shards on the GPUs) and block fetch
mutable struct GramHandle
    ptr::Ptr{Cvoid}
    n::Int
end
function pl_gram_resident(X::Matrix{Float64}, kid::Cint, α::Cint, kind::Cint, cinv, ℓ::Float64, β::Float64)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pl_gram_resident, libplb200), Cint, (Cint, Ptr{Cdouble}, Int64, Int64, Int64, Cint, Cint, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Ptr{Cvoid}}),
                kid, X, size(X, 2), size(X, 1), size(X, 1), α, kind, cinv, ℓ, β, h), "pl_gram_resident")
    H = GramHandle(h[], size(X, 2))
    finalizer(x -> (x.ptr == C_NULL || ccall((:pl_gram_free, libplb200), Cint, (Ptr{Cvoid},), x.ptr); x.ptr = C_NULL), H)
    H
end
# K[i, j] for i in rows, j in cols (1-based unit ranges), any block of the full symmetric matrix; result is length(rows) x length(cols)
function fetch_block(H::GramHandle, rows::UnitRange{Int}, cols::UnitRange{Int})
    out = zeros(length(cols), length(rows))                    # row-major (rows x cols) of the library == column-major cols x rows
    check(ccall((:pl_gram_fetch_block, libplb200), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Ptr{Cdouble}, Int64),
                H.ptr, first(rows) - 1, last(rows), first(cols) - 1, last(cols), out, length(cols)), "pl_gram_fetch_block")
    permutedims(out)
end

# ---- normal equations (linear-learn.jl: every A'WA / A'Wb site) -------------------------------------------------------------------
# A: K x R column-major (rows of the design matrix as columns). w: per-row weights or nothing (then the two-segment form we / wf).
function pl_normal_eq(A::Matrix{Float64}, b::Union{Vector{Float64},Nothing}, ne::Integer, we::Real, wf::Real, int::Bool, λ::Real; w = nothing)
    K, R = size(A); n = K + (int ? 1 : 0)
    AtWA = zeros(n, n); AtWb = zeros(n)
    check(ccall((:pl_normal_eq, libplb200), Cint,
                (Ptr{Cdouble}, Int64, Int64, Int64, Ptr{Cdouble}, Int64, Cdouble, Cdouble, Ptr{Cdouble}, Cint, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}),
                A, R, K, K, w === nothing ? C_NULL : w, ne, Float64(we), Float64(wf), b === nothing ? C_NULL : b, Cint(int), Float64(λ), AtWA,
                b === nothing ? C_NULL : AtWb), "pl_normal_eq")
    AtWA, AtWb
end

# the same over a LIST of blocks as they lie in memory: lp.dB[i] (K x 3natoms column-major) IS 3natoms rows of K doubles — no hcat copy
# of the 11.5 GB design matrix; the library streams row chunks (over every GPU's own PCIe link after init!(devices = "all"))
function pl_normal_eq_blocks(blocks::Vector{Matrix{Float64}}, b::Vector{Float64}, ne::Integer, we::Real, wf::Real, int::Bool, λ::Real)
    K = size(blocks[1], 1); n = K + (int ? 1 : 0)
    ptrs = Ptr{Cdouble}[pointer(B) for B in blocks]; rows = Int64[size(B, 2) for B in blocks]
    AtWA = zeros(n, n); AtWb = zeros(n)
    GC.@preserve blocks check(ccall((:pl_normal_eq_blocks, libplb200), Cint,
                (Ptr{Ptr{Cdouble}}, Ptr{Int64}, Int64, Int64, Ptr{Cdouble}, Int64, Cdouble, Cdouble, Ptr{Cdouble}, Cint, Cdouble, Int64, Ptr{Cdouble}, Ptr{Cdouble}),
                ptrs, rows, length(blocks), K, C_NULL, ne, Float64(we), Float64(wf), b, Cint(int), Float64(λ), 0, AtWA, AtWb), "pl_normal_eq_blocks")
    AtWA, AtWb
end

# out-of-core running sums of ooc_learn! (linear-learn.jl:355-356) resident on the device
neq_begin(K::Integer) = check(ccall((:pl_neq_begin, libplb200), Cint, (Int64,), K), "pl_neq_begin")
function neq_add(At::Matrix{Float64}, b::Vector{Float64}, w::Vector{Float64})                                  # At: K x rows
    K, R = size(At)
    check(ccall((:pl_neq_add, libplb200), Cint, (Ptr{Cdouble}, Int64, Int64, Ptr{Cdouble}, Int64, Cdouble, Cdouble, Ptr{Cdouble}),
                At, R, K, w, 0, 1.0, 1.0, b), "pl_neq_add")
end
function neq_finish(K::Integer; int::Bool = false, λ::Real = 0.0)
    n = K + (int ? 1 : 0); AtWA = zeros(n, n); AtWb = zeros(n)
    check(ccall((:pl_neq_finish, libplb200), Cint, (Cint, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}), Cint(int), Float64(λ), AtWA, AtWb), "pl_neq_finish")
    AtWA, AtWb
end

# ---- covariance / second moment (dimension_reduction.jl:11-12, pca_state.jl:34-37) ---------------------------------------------------
function pl_cov(X::Matrix{Float64}, m::Vector{Float64}, mean_given::Bool, center::Bool)                        # X: K x n
    K, n = size(X); C = zeros(K, K)
    check(ccall((:pl_cov, libplb200), Cint, (Ptr{Cdouble}, Int64, Int64, Int64, Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}),
                X, n, K, K, m, Cint(mean_given), Cint(center), C), "pl_cov")
    C
end

# eigen(Symmetric(Q)) on the device (the library's own solver up to n = 2048, cuSOLVER above; OPT-IN — the default keeps LAPACK like the reference): row k of the
# library's V is column k of Julia's `vectors`
function pl_eigh(A::Matrix{Float64}, uplo::Char)
    n = size(A, 1); W = zeros(n); V = zeros(n, n)
    check(ccall((:pl_eigh, libplb200), Cint, (Ptr{Cdouble}, Int64, Int64, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Int64),
                A, n, n, uplo == 'U' ? Cint(2) : Cint(1), W, V, n), "pl_eigh")
    W, V
end

# ---- features (features.jl:19-50) --------------------------------------------------------------------------------------------------
function pl_global_feature(L::Matrix{Float64}, offsets::Vector{Int64}, mean::Bool)                             # L: d x total atoms
    d = size(L, 1); N = length(offsets) - 1; X = zeros(d, N)
    check(ccall((:pl_global_feature, libplb200), Cint, (Ptr{Cdouble}, Int64, Ptr{Int64}, Int64, Int64, Cint, Ptr{Cdouble}, Int64),
                L, d, offsets, N, d, Cint(mean), X, d), "pl_global_feature")
    X
end
function pl_correlation_feature(L::Matrix{Float64}, offsets::Vector{Int64})
    d = size(L, 1); N = length(offsets) - 1; C = zeros(d, d, N)
    check(ccall((:pl_correlation_feature, libplb200), Cint, (Ptr{Cdouble}, Int64, Ptr{Int64}, Int64, Int64, Ptr{Cdouble}),
                L, d, offsets, N, d, C), "pl_correlation_feature")
    C
end
function pl_ksd_rbf(X::Matrix{Float64}, S::Matrix{Float64}, kind::Cint, cinv, ℓ::Float64, β::Float64)
    d, n = size(X); out = zeros(1)
    check(ccall((:pl_ksd_rbf, libplb200), Cint, (Ptr{Cdouble}, Ptr{Cdouble}, Int64, Int64, Int64, Int64, Cint, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cdouble}),
                X, S, n, d, d, d, kind, cinv, ℓ, β, out), "pl_ksd_rbf")
    out[1]
end

# ---- after the fit: predictions and metrics (Data/utils.jl:42-65, Metrics/metrics.jl:11-53) -----------------------------------------
function pl_predict(A::Matrix{Float64}, β::Vector{Float64}, β0::Float64, ne::Integer)                          # A: K x R
    K, R = size(A); y = zeros(R)
    check(ccall((:pl_predict, libplb200), Cint, (Ptr{Cdouble}, Int64, Int64, Int64, Ptr{Cdouble}, Cdouble, Int64, Ptr{Cdouble}),
                A, R, K, K, β, β0, ne, y), "pl_predict")
    y
end
function pl_metrics(x_pred::Vector{Float64}, x::Vector{Float64})
    out = zeros(4)
    check(ccall((:pl_metrics, libplb200), Cint, (Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}), x_pred, x, length(x), out), "pl_metrics")
    (mae = out[1], rmse = out[2], rsq = out[3], mean_cos = out[4])
end

# ---- L-ensemble on the device (optional, outside the contracted path: Determinantal.jl keeps sampling) --------------------------------
function pl_ell_ensemble(X::Matrix{Float64}, kid::Cint, α::Cint, kind::Cint, cinv, ℓ::Float64, β::Float64)
    d, n = size(X); λ = zeros(n); K = zeros(n, n)
    check(ccall((:pl_ell_ensemble, libplb200), Cint,
                (Cint, Ptr{Cdouble}, Int64, Int64, Int64, Cint, Cint, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Int64),
                kid, X, n, d, d, α, kind, cinv, ℓ, β, λ, K, n), "pl_ell_ensemble")
    λ, K
end
ell_inclusion_prob(α::Real, n::Integer) = (p = zeros(n);
    check(ccall((:pl_ell_inclusion_prob, libplb200), Cint, (Cdouble, Int64, Ptr{Cdouble}), Float64(α), n, p), "pl_ell_inclusion_prob"); p)
ell_eigvecs(idx::Vector{Int64}, n::Integer) = (V = zeros(n, length(idx));                # 0-based eigenvector indices; column j = U[:, idx[j]+1]
    check(ccall((:pl_ell_eigvecs, libplb200), Cint, (Ptr{Int64}, Int64, Int64, Ptr{Cdouble}, Int64), idx, length(idx), n, V, n), "pl_ell_eigvecs"); V)

# ---- DBSCAN's distance matrices (dbscan.jl:132-170): positions 3 x natoms x n column-major == n x natoms x 3 row-major ---------------
function distance_matrix_gpu(P::Array{Float64,3}; box_lengths::Union{Nothing,Vector{Float64}} = nothing)
    natoms, n = size(P, 2), size(P, 3); D = zeros(n, n)
    if box_lengths === nothing
        check(ccall((:pl_rmsd_kabsch, libplb200), Cint, (Ptr{Cdouble}, Int64, Int64, Ptr{Cdouble}, Int64), P, n, natoms, D, n), "pl_rmsd_kabsch")
    else
        check(ccall((:pl_rmsd_periodic, libplb200), Cint, (Ptr{Cdouble}, Int64, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Int64), P, n, natoms, box_lengths, D, n), "pl_rmsd_periodic")
    end
    D
end
