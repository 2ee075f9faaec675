# kernels.jl — KernelMatrix on the B200 (reference: src/Kernels/kernels.jl:211-269, features.jl:19-50, divergences.jl:27-49).

# Vector{Vector{Float64}} -> d x n column-major == the library's row-major n x d (zero transposition)
pack(F::Vector{Vector{Float64}}) = reduce(hcat, F)
pack(F::Vector{<:Vector{<:Real}}) = Matrix{Float64}(reduce(hcat, F))

# (cinv_kind, pointer-able argument) of a Euclidean distance (distances.jl:39-51). The second value is passed to ccall directly, which
# keeps it rooted for the duration of the call.
function cinv_args(e::Euclidean)
    C = e.Cinv
    if C isa Diagonal
        all(==(1.0), C.diag) && return (PL_CINV_IDENTITY, C_NULL)
        return (PL_CINV_DIAGONAL, Vector{Float64}(C.diag))
    end
    (PL_CINV_FULL, Matrix{Float64}(C))
end

supported(k) = k isa DotProduct || ((k isa RBF || k isa PotentialLearning.InverseMultiquadric) && k.d isa Euclidean)

"""
    kernel_matrix(F, k) -> Symmetric{Float64,Matrix{Float64}}

`KernelMatrix(F, k)` (kernels.jl:211-223) for `DotProduct`, `RBF(Euclidean)` and `InverseMultiquadric(Euclidean)`: only the `j >= i`
triangle is computed and stored, over `zeros(n, n)`, and wrapped as `Symmetric(K)` (uplo = :U) exactly like the reference.
`RBF.α` is never applied (kernels.jl:73-74); `K[i, i] == β` exactly.
"""
function kernel_matrix(F::Vector{<:Vector{<:Real}}, k)
    X = pack(F); n = size(X, 2)
    K = result_zeros(n, n)                                                             # kernels.jl:216
    if k isa DotProduct
        pl_gram_dot!(K, X, k.α)                                                        # kernels.jl:35
    elseif k isa RBF
        kind, cinv = cinv_args(k.d)
        pl_gram_rbf!(K, X, kind, cinv, Float64(k.ℓ), Float64(k.β))                      # kernels.jl:73-74, distances.jl:58-60
    else
        kind, cinv = cinv_args(k.d)
        pl_gram_imq!(K, X, kind, cinv, Float64(k.c2), Float64(k.ℓ))                     # kernels.jl:156-164
    end
    Symmetric(K)                                                                       # kernels.jl:222
end

"""
    kernel_matrix(F1, F2, k) -> Matrix{Float64}  (m x n, kernels.jl:229-244)
"""
function kernel_matrix(F1::Vector{<:Vector{<:Real}}, F2::Vector{<:Vector{<:Real}}, k)
    X1 = pack(F1); X2 = pack(F2)
    K = zeros(size(X1, 2), size(X2, 2))
    if k isa DotProduct
        pl_gram_cross!(K, X1, X2, PL_KERNEL_DOT, Cint(k.α), PL_CINV_IDENTITY, C_NULL, 1.0, 1.0)
    else
        kind, cinv = cinv_args(k.d)
        k isa RBF ? pl_gram_cross!(K, X1, X2, PL_KERNEL_RBF, Cint(0), kind, cinv, Float64(k.ℓ), Float64(k.β)) :
                    pl_gram_cross!(K, X1, X2, PL_KERNEL_IMQ, Cint(0), kind, cinv, Float64(k.ℓ), Float64(k.c2))
    end
    K
end

"""
    gram_resident(F, k) -> GramHandle ;  fetch_block(H, rows, cols)

The kernel matrix computed and LEFT on the GPUs as tile-packed shards (N = 200 000 has a 160 GB upper triangle: it cannot be a host
`Matrix`). `fetch_block` gathers any block of the full symmetric matrix.
"""
function gram_resident(F::Vector{<:Vector{<:Real}}, k)
    X = pack(F)
    if k isa DotProduct
        return pl_gram_resident(X, PL_KERNEL_DOT, Cint(k.α), PL_CINV_IDENTITY, C_NULL, 1.0, 1.0)
    end
    kind, cinv = cinv_args(k.d)
    k isa RBF ? pl_gram_resident(X, PL_KERNEL_RBF, Cint(0), kind, cinv, Float64(k.ℓ), Float64(k.β)) :
                pl_gram_resident(X, PL_KERNEL_IMQ, Cint(0), kind, cinv, Float64(k.ℓ), Float64(k.c2))
end

# ---- features of a whole DataSet in one segmented reduction (features.jl:19-34, 48-50, 76-78) ---------------------------------------
function stacked_locals(ds)
    locals = [reduce(hcat, PotentialLearning.get_values(PotentialLearning.get_local_descriptors(c))) for c in ds]   # d x natoms_i each
    L = Matrix{Float64}(reduce(hcat, locals))
    L, Int64[0; cumsum(size.(locals, 2))]
end
function global_features(ds, mean::Bool)                       # Julia's summation order inside the kernel: bit-identical to sum / mean
    L, offsets = stacked_locals(ds)
    X = pl_global_feature(L, offsets, mean)
    [X[:, i] for i in 1:size(X, 2)]
end
function correlation_features(ds)
    L, offsets = stacked_locals(ds)
    C = pl_correlation_feature(L, offsets)
    [Symmetric(C[:, :, i]) for i in 1:size(C, 3)]
end

# ---- KernelSteinDiscrepancy with an RBF(Euclidean) kernel (divergences.jl:27-49) ----------------------------------------------------
function ksd_rbf(x::Vector{Vector{Float64}}, div)
    k = div.kernel
    X = pack(x); S = pack(Vector{Float64}.(div.score.(x)))
    kind, cinv = cinv_args(k.d)
    pl_ksd_rbf(X, S, kind, cinv, Float64(k.ℓ), Float64(k.β))
end

# ---- EllEnsemble(KernelMatrix(F, k)) with K and U resident on the device (optional; Determinantal.jl keeps `sample`) ------------------
function ell_ensemble_gpu(F::Vector{<:Vector{<:Real}}, k)
    X = pack(F)
    λ, K = k isa DotProduct ? pl_ell_ensemble(X, PL_KERNEL_DOT, Cint(k.α), PL_CINV_IDENTITY, C_NULL, 1.0, 1.0) :
                              pl_ell_ensemble(X, PL_KERNEL_RBF, Cint(0), cinv_args(k.d)..., Float64(k.ℓ), Float64(k.β))
    λ, Symmetric(K)
end
