# PotentialLearningB200 — the Julia side of the drop-in. `using PotentialLearning, PotentialLearningB200` re-points the descriptor
# linear-algebra hot path of PotentialLearning.jl (v0.2.7) at libplb200.so (hand-written sm_100a CUDA behind the C ABI of
# include/plb200.h). There are no CUDA.jl kernels and there is no CPU fallback: a non-zero status raises.
#
# Layout of this package
#   src/lib.jl        one thin Julia function per C entry point (the only place with `ccall`)
#   src/kernels.jl    KernelMatrix forms, features, KSD                (src/Kernels of the reference)
#   src/learning.jl   normal-equation assembly for every learn! form   (src/Learning/linear-learn.jl)
#   src/dimred.jl     covariance / second moment, PCAState fit!        (src/DimensionReduction)
#   src/overrides.jl  the methods that are evaluated INTO PotentialLearning by activate!()
#
# Why activate!() and not plain method definitions: several of the reference's methods have to be replaced with an IDENTICAL signature
# (learn!(::UnivariateLinearProblem, ::Vector, ::Bool), compute_eigen, fit!(::DataSet, ::PCAState), ...). Overwriting another module's
# method while this package precompiles is an error on Julia >= 1.10 (the reference's floor, Project.toml:32). The overrides are
# therefore an expression (src/overrides.jl) that activate!() evaluates into the already loaded PotentialLearning module at RUN time —
# never during precompilation (`jl_generating_output`). __init__ calls it unless PLB200_AUTO_ACTIVATE=0; every implementation is also
# callable directly (PotentialLearningB200.kernel_matrix, .normal_equations, .covariance, ...) without activating anything.
#
# NOTE: Julia is not installed in the image this repository is built and benchmarked in, so this package has not been executed there.
# tests/test_julia_shim_static.py checks every `ccall` against include/plb200.h (names, argument types, order), that no method with a
# reference-identical signature is defined outside the activation block, and tests/c_abi_driver.c drives the same argument lists from
# plain C. test/runtests.jl is the Julia-side suite for a machine that has Julia + a B200.
module PotentialLearningB200

using LinearAlgebra
using Statistics
using PotentialLearning

export activate!, init!, device_count, kernel_matrix, normal_equations, covariance, eigen_gpu, ell_ensemble_gpu, gram_resident, fetch_block

include("lib.jl")
include("kernels.jl")
include("learning.jl")
include("dimred.jl")
include("overrides.jl")

const ACTIVE = Ref(false)

"""
    activate!()

Evaluate the overriding methods (src/overrides.jl) into the loaded `PotentialLearning` module. Idempotent. Run-time only: a no-op
while any package is being precompiled.
"""
function activate!()
    ccall(:jl_generating_output, Cint, ()) == 1 && return false
    ACTIVE[] && return true
    Core.eval(PotentialLearning, OVERRIDES)
    ACTIVE[] = true
end

"""
    init!(; devices = nothing)

Initialise libplb200. `devices = nothing` reads `PLB200_DEVICES` ("all", or a list such as "0,1,2,3"; default: one GPU, `PLB200_DEVICE`
or the current device). With more than one device ONE Julia process drives them all (`pl_init_all`: a context and an NCCL communicator
per GPU inside the library); `learn!`, `fit!` and `KernelMatrix` then shard large inputs over the GPUs transparently.
"""
function init!(; devices = nothing)
    spec = devices === nothing ? get(ENV, "PLB200_DEVICES", "") : devices
    if spec isa AbstractString
        s = strip(spec)
        if isempty(s)
            return pl_init(parse(Cint, get(ENV, "PLB200_DEVICE", "-1")))
        elseif lowercase(s) == "all"
            return pl_init_all(Cint[])
        else
            return pl_init_all(Cint[parse(Cint, t) for t in split(s, ",")])
        end
    end
    pl_init_all(Cint.(collect(spec)))
end

function __init__()
    ccall(:jl_generating_output, Cint, ()) == 1 && return       # precompiling some package: touch nothing
    init!()                                                      # errors without an sm_100 GPU: there is no CPU path
    atexit(pl_finalize)
    get(ENV, "PLB200_AUTO_ACTIVATE", "1") == "0" || activate!()
    nothing
end

end # module
