# overrides.jl — the methods activate!() evaluates INTO the loaded PotentialLearning module (run time only, never while precompiling).
# Inside the block every unqualified name is resolved in PotentialLearning's own namespace (Optimization, Flux, MvNormal, ProgressBar,
# get_values, ... are the reference's own imports); `B200` is this package. Each method keeps the reference's signature, return value and
# host-side steps; only the dense contractions go to the GPU. Methods whose signature is IDENTICAL to the reference's (they overwrite)
# are marked (=); the others are more specific than the reference's generic ones and fall back to it with `invoke` for kernels /
# distances the library does not cover (Forstner, ...).

const OVERRIDES = quote
    const B200 = $(@__MODULE__)

    # ---- KernelMatrix (kernels.jl:211-244). kDPP(ds, f, k) / kDPP(features, k) (dpp.jl:18-44) need nothing: they call these. ----------
    function KernelMatrix(F::Vector{Vector{Float64}}, k::Union{DotProduct,RBF,InverseMultiquadric})
        B200.supported(k) || return invoke(KernelMatrix, Tuple{Union{Vector{Vector{Float64}},Vector{Symmetric{Float64,Matrix{Float64}}}},Kernel}, F, k)
        B200.kernel_matrix(F, k)
    end
    function KernelMatrix(F1::Vector{Vector{Float64}}, F2::Vector{Vector{Float64}}, k::Union{DotProduct,RBF,InverseMultiquadric})
        B200.supported(k) || return invoke(KernelMatrix, Tuple{Union{Vector{Vector{Float64}},Vector{Symmetric{Float64,Matrix{Float64}}}},
                                                               Union{Vector{Vector{Float64}},Vector{Symmetric{Float64,Matrix{Float64}}}},Kernel}, F1, F2, k)
        B200.kernel_matrix(F1, F2, k)
    end

    # ---- features of a whole DataSet (features.jl:76-78 over :19-34, :48-50); GlobalMean / GlobalSum ignore dt like the reference --------
    compute_features(ds::DataSet, f::GlobalMean; dt = LocalDescriptors) = B200.global_features(ds, true)
    compute_features(ds::DataSet, f::GlobalSum; dt = LocalDescriptors) = B200.global_features(ds, false)
    function compute_features(ds::DataSet, f::CorrelationMatrix; dt = LocalDescriptors)
        dt === LocalDescriptors || return compute_feature.(ds, (f,); dt = dt)
        B200.correlation_features(ds)
    end

    # ---- KernelSteinDiscrepancy with RBF(Euclidean) (divergences.jl:27-49) ------------------------------------------------------------
    function compute_divergence(x::Vector{Vector{Float64}}, div::KernelSteinDiscrepancy)
        k = div.kernel
        (k isa RBF && k.d isa Euclidean) || return invoke(compute_divergence, Tuple{Vector{<:Union{Real,Vector{<:Real}}},KernelSteinDiscrepancy}, x, div)
        B200.ksd_rbf(x, div)
    end

    # ---- (=) learn!(lp::UnivariateLinearProblem, α) — OLS (linear-learn.jl:11-23): the two sums on the GPU, pinv / std on the host -----
    function learn!(lp::UnivariateLinearProblem, α::Real)
        AtA, Atb = B200.outer_sums(lp.iv_data, lp.dv_data)                              # :16-17
        Q = pinv(AtA, α)
        lp.β .= Q * Atb
        lp.σ .= std(Atb - AtA * lp.β)
        lp.Σ .= Symmetric(lp.σ[1]^2 * Q)
    end

    # ---- (=) learn!(lp::CovariateLinearProblem, α) — MLE (linear-learn.jl:37-69): the four sums on the GPU, Adam on the host ------------
    function learn!(lp::CovariateLinearProblem, α::Real)
        AtAe, Atbe = B200.outer_sums(lp.B, lp.e)                                        # :45-46
        AtAf, Atbf = B200.outer_sums(lp.dB, lp.f)                                       # :48-49
        f(x, p) =
            -logpdf(MvNormal(p[1] * x[3:end], exp(x[1]) + p[5]), p[2]) -
            logpdf(MvNormal(p[3] * x[3:end], exp(x[2]) + p[5]), p[4])
        g = Optimization.OptimizationFunction(f, Optimization.AutoForwardDiff())
        x0 = [log(lp.σe[1]), log(lp.σf[1]), lp.β...]
        p = [AtAe, Atbe, AtAf, Atbf, α]
        prob = Optimization.OptimizationProblem(g, x0, p)
        sol = Optimization.solve(prob, OptimizationOptimisers.Adam(), maxiters = 1000)
        lp.σe .= exp(sol.u[1])
        lp.σf .= exp(sol.u[2])
        lp.β .= sol.u[3:end]
        Q = pinv(Symmetric(lp.σe[1]^2 * pinv(Symmetric(AtAe), α) + lp.σf[1]^2 * pinv(Symmetric(AtAf), α)))
        lp.Σ .= Symmetric(Q)
    end

    # ---- (=) minibatch learn!(lp, ss, α) (linear-learn.jl:82-165): the per-batch sums and the final full sums on the GPU ---------------
    function learn!(lp::UnivariateLinearProblem, ss::SubsetSelector, α::Real; num_steps = 100, opt = Flux.Optimise.Adam())
        params = [log.(lp.σ); lp.β]
        f(x, p) = -logpdf(MvNormal(p[1] * x[2:end], exp(x[1])), p[2])
        for step = 1:num_steps
            inds = get_random_subset(ss)
            p = B200.outer_sums(lp.iv_data[inds], lp.dv_data[inds])                     # :94-95
            if step % (num_steps ÷ 10) == 0
                err = @sprintf("%1.3e", f(params, p))
                println("Iteration #$(step): \t log(p(x)) = $err")
            end
            grads = Flux.gradient(x -> f(x, p), params)[1]
            Flux.Optimise.update!(opt, params, grads)
        end
        lp.σ .= exp(params[1])
        lp.β .= params[2:end]
        AtA = B200.outer_sum(lp.iv_data)                                                # :107
        lp.Σ .= Symmetric(lp.σ[1]^2 * pinv(AtA, α))
    end
    function learn!(lp::CovariateLinearProblem, ss::SubsetSelector, α::Real; num_steps = 100, opt = Flux.Optimise.Adam())
        params = [log.(lp.σe); log.(lp.σf); lp.β]
        f(x, p) =
            -logpdf(MvNormal(p[1] * x[3:end], exp(x[1]) + p[5]), p[2]) -
            logpdf(MvNormal(p[3] * x[3:end], exp(x[2]) + p[5]), p[4])
        for step = 1:num_steps
            inds = get_random_subset(ss)
            AtAe, Atbe = B200.outer_sums(lp.B[inds], lp.e[inds])                        # :139-140
            AtAf, Atbf = B200.outer_sums(lp.dB[inds], lp.f[inds])                       # :142-143
            p = (AtAe, Atbe, AtAf, Atbf)   # four parameters, as in the reference (:145; f reads p[5] — the reference carries the same TODO)
            if step % (num_steps ÷ 10) == 0
                err = @sprintf("%1.3e", f(params, p))
                println("Iteration #$(step): \t Batch log(p(x)) = $err")
            end
            grads = Flux.gradient(x -> f(x, p), params)
            Flux.Optimise.update!(opt, params, grads)
        end
        lp.σe .= exp(params[1])
        lp.σf .= exp(params[2])
        lp.β .= params[3:end]
        AtAe = B200.outer_sum(lp.B)                                                     # :157
        AtAf = B200.outer_sum(lp.dB)                                                    # :158
        Q = pinv(Symmetric(lp.σe[1]^2 * pinv(Symmetric(AtAe), α) + lp.σf[1]^2 * pinv(Symmetric(AtAf), α)))
        lp.Σ .= Symmetric(Q)
    end

    # ---- (=) WLS learn!(lp, ws, int[; λ]) (linear-learn.jl:178-268): A'QA (+ λI), A'Qb on the GPU, try-`\`/catch-pinv on the host -------
    function learn!(lp::UnivariateLinearProblem, ws::Vector, int::Bool)
        A = B200.pack_rows(lp.iv_data); b = Float64.(reduce(vcat, lp.dv_data))
        M, v = B200.normal_equations(A, b, size(A, 2), ws[1], ws[1]; int = int)         # Q = ws[1] * I (:196); intercept = ones (:188)
        βs = B200.solve_normal(M, v)
        if int
            lp.β0 .= βs[1]
            lp.β .= βs[2:end]
        else
            lp.β .= βs
        end
    end
    function learn!(lp::CovariateLinearProblem, ws::Vector, int::Bool; λ::Real = 0.0)
        M, v = B200.wls_matrices(lp, ws, int, λ)                                        # :232-253
        βs = B200.solve_normal(M, v)
        if int
            lp.β0 .= βs[1]
            lp.β .= βs[2:end]
        else
            lp.β .= βs
        end
    end

    # ---- (=) ooc_learn! (linear-learn.jl:301-391): the running sums AtWA .+= A'WA, AtWb .+= A'Wb stay on the device -------------------
    function ooc_learn!(lb::InteratomicPotentials.LinearBasisPotential, ds_train::PotentialLearning.DataSet;
                        ws = [30.0, 1.0], symmetrize::Bool = true, λ::Union{Real,Nothing} = 0.01, reg_style::Symbol = :default,
                        AtWA = nothing, AtWb = nothing, pbar = true, eweight_normalized = :squared)
        basis_size = length(lb.basis)
        if isnothing(AtWA) || isnothing(AtWb)
            B200.neq_begin(basis_size)
            iter = pbar ? ProgressBar(ds_train) : ds_train
            for config in iter
                ref_energy = get_values(get_energy(config))
                ref_forces = reduce(vcat, get_values(get_forces(config)))
                sys = get_system(config)
                natoms = length(sys)
                global_descrs = sum(compute_local_descriptors(sys, lb.basis))                        # K-vector (:334)
                force_descrs = stack(reduce(vcat, compute_force_descriptors(sys, lb.basis)))          # K x 3natoms (:335, untransposed)
                we_norm = isnothing(eweight_normalized) ? 1.0 :
                          eweight_normalized == :standard ? 1 / natoms :
                          eweight_normalized == :squared ? 1 / natoms^2 :
                          error("eweight_normalized can only be nothing, :standard, or :squared")
                At = Matrix{Float64}(hcat(global_descrs, force_descrs))                              # K x (1 + 3natoms): rows of A as columns
                b = Float64[ref_energy; ref_forces]
                w = Float64[we_norm * ws[1]; ws[2] * ones(length(ref_forces))]                       # diag(W) (:351-352)
                B200.neq_add(At, b, w)                                                               # :355-356
            end
            AtWA, AtWb = B200.neq_finish(basis_size)
        else
            AtWA = deepcopy(AtWA)
            AtWb = deepcopy(AtWb)
        end
        if symmetrize
            AtWA = Symmetric(AtWA)
        end
        if !isnothing(λ)
            if reg_style == :default
                AtWA += λ * Diagonal(ones(size(AtWA)[1]))
            end
            if reg_style == :scale_thresh || reg_style == :scale
                for i in 1:size(AtWA, 1)
                    reg_elem = AtWA[i, i] * (1 + λ)
                    if reg_style == :scale_thresh
                        reg_elem = max(reg_elem, λ)
                    end
                    AtWA[i, i] = reg_elem
                end
            end
        end
        β = AtWA \ AtWb
        println("condition number of AtWA: $(cond(AtWA))")
        lb.β .= β
        AtWA, AtWb
    end

    # ---- (=) compute_eigen (dimension_reduction.jl:11-14): Q on the GPU, eigen(Symmetric(Q)) in LAPACK as in the reference ---------------
    function compute_eigen(d::Vector{T}) where {T<:Vector{<:Real}}
        Q, _ = B200.covariance(d)
        B200.eigen_of(Q)
    end

    # ---- (=) fit!(ds, pca::PCAState) (pca_state.jl:20-40): sums, mean, centred covariance in one pass on the GPU ---------------------------
    fit!(ds::DataSet, pca::PCAState) = B200.fit_pcastate!(ds, pca)
end
