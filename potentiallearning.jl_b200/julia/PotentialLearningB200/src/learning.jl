# learning.jl — every A'WA / A'Wb assembly site of src/Learning/linear-learn.jl on the B200. The solves (`\`, pinv), the Adam / MvNormal
# loops and everything else stay in Julia exactly as the reference has them (src/overrides.jl).

# rows of the design matrix as the library wants them: hcat of K-vectors / K x m blocks == column-major K x R == row-major R x K
pack_rows(blocks) = Matrix{Float64}(reduce(hcat, blocks))
as_block(M::Matrix{Float64}) = M                                # lp.dB[i] is handed over as it lies in memory: NO copy
as_block(M::AbstractMatrix) = Matrix{Float64}(M)

"""
    normal_equations(A, b, ne, we, wf; int = false, λ = 0.0, w = nothing) -> (AtWA, AtWb)

`A` is K x R (one design-matrix row per column, rows ordered [ne energy rows ; force rows] like `A = [B_train; dB_train]`,
linear-learn.jl:238-244). Weights: `w` per row, or `we` for the first `ne` rows and `wf` for the rest (`Q`, :246-249).
`int` prepends the intercept column `[1_N; 0_M]` (:239-241); `λ` lands on every diagonal entry, intercept included (:253).
"""
normal_equations(A::Matrix{Float64}, b, ne::Integer, we::Real, wf::Real; int::Bool = false, λ::Real = 0.0, w = nothing) =
    pl_normal_eq(A, b === nothing ? nothing : Vector{Float64}(b), ne, we, wf, int, λ; w = w)

# sum(v * v' for v in vs), sum(v * y for (v, y) in zip(vs, ys)) — the OLS / MLE sums (linear-learn.jl:16-17, 45-49, 94-95, 139-143, 157-158).
# vs: K-vectors (energies) or K x m matrices (forces; ys then m-vectors).
function outer_sums(vs, ys)
    A = pack_rows(vs); b = Float64.(reduce(vcat, ys))
    normal_equations(A, b, 0, 1.0, 1.0)
end
outer_sum(vs) = first(pl_normal_eq(pack_rows(vs), nothing, 0, 1.0, 1.0, false, 0.0))

# WLS of a CovariateLinearProblem (linear-learn.jl:226-268): lp.B packed once (N x K is small), every lp.dB[i] passed as it lies in memory
function wls_matrices(lp, ws::Vector, int::Bool, λ::Real)
    blocks = Matrix{Float64}[pack_rows(lp.B)]
    for dBi in lp.dB
        push!(blocks, as_block(dBi))
    end
    b = Float64.(vcat(lp.e, reduce(vcat, lp.f)))
    pl_normal_eq_blocks(blocks, b, length(lp.e), ws[1], ws[2], int, λ)
end

# (AtWA, AtWb) of exactly what assemble_matrices(lp, ws) (linear-learn.jl:286-299) describes: A'*W*A, A'*W*b without materialising A or W
normal_matrices(lp, ws::Vector) = wls_matrices(lp, ws, false, 0.0)

# the try-`\` / catch-pinv solve of linear-learn.jl:199-205, 252-258 on GPU-assembled matrices
function solve_normal(M, v)
    try
        M \ v
    catch e
        println(e)
        println("Linear system will be solved using pinv.")
        pinv(M) * v
    end
end
