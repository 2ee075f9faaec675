"""Static checks of the Julia package (potentiallearning.jl_b200/julia/PotentialLearningB200) against include/plb200.h and against the
reference's method table. Julia is not installed in this image, so the package cannot be executed here — but
  * every `ccall` must name an exported function and pass exactly the argument types the header declares, in order;
  * all `ccall`s live in src/lib.jl;
  * NO method with a signature identical to one of the reference's may be defined at package level (that is method overwriting from
    another module: an error while precompiling on Julia >= 1.10) — such methods may only appear inside the `OVERRIDES` expression that
    activate!() evaluates into PotentialLearning at run time, and activation / initialisation are skipped while precompiling;
  * the package has the loadable layout (Project.toml with name / uuid / deps, src/<Name>.jl, test/runtests.jl)."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "plb200.h")
PKG = os.path.join(ROOT, "potentiallearning.jl_b200", "julia", "PotentialLearningB200")
SRC = os.path.join(PKG, "src")

C2JL = {
    "const double*": "Ptr{Cdouble}", "double*": "Ptr{Cdouble}", "const int64_t*": "Ptr{Int64}", "int64_t*": "Ptr{Int64}",
    "const double* const*": "Ptr{Ptr{Cdouble}}", "int64_t": "Int64", "int": "Cint", "double": "Cdouble", "void*": "Ptr{Cvoid}",
    "const int*": "Ptr{Cint}", "int*": "Ptr{Cint}", "const void*": "Ptr{Cvoid}", "void**": "Ptr{Ptr{Cvoid}}",
}
RET2JL = {"int": "Cint", "void": "Cvoid", "const char*": "Cstring", "int64_t": "Int64", "void*": "Ptr{Cvoid}"}


def header_prototypes():
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    txt = "\n".join(l for l in txt.splitlines() if not l.lstrip().startswith("#"))  # preprocessor lines
    protos = {}
    for m in re.finditer(r"((?:const\s+)?[A-Za-z_]\w*\s*\**)\s*\b(pl_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", txt, flags=re.S):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3)
        types = []
        for p in [q.strip() for q in params.replace("\n", " ").split(",") if q.strip() and q.strip() != "void"]:
            t = re.sub(r"\s*\b[A-Za-z_]\w*$", "", p).strip()  # drop the parameter name
            t = re.sub(r"\s*\*", "*", t)
            t = re.sub(r"\s+", " ", t).replace("* const*", "* const*")
            types.append(t)
        protos[name] = (re.sub(r"\s*\*", "*", ret), types)
    return protos


def sources():
    return {os.path.basename(p): open(p, encoding="utf-8").read() for p in sorted(glob.glob(os.path.join(SRC, "*.jl")))}


def ccalls(src):
    calls = []
    for m in re.finditer(r"ccall\(\(:(pl_\w+),\s*libplb200\),\s*([\w{}]+),\s*\(([^)]*)\)", src, flags=re.S):
        args = [a.strip() for a in m.group(3).replace("\n", " ").split(",") if a.strip()]
        calls.append((m.group(1), m.group(2), args))
    return calls


def test_package_layout():
    toml = open(os.path.join(PKG, "Project.toml")).read()
    assert re.search(r'^name = "PotentialLearningB200"$', toml, flags=re.M) and re.search(r'^uuid = "[0-9a-f-]{36}"$', toml, flags=re.M)
    assert 'PotentialLearning = "82b0a93c-c2e3-44bc-a418-f0f89b0ae5c2"' in toml, "depends on the reference by its registered uuid (reference Project.toml:2)"
    assert os.path.exists(os.path.join(SRC, "PotentialLearningB200.jl")) and os.path.exists(os.path.join(PKG, "test", "runtests.jl"))
    main = sources()["PotentialLearningB200.jl"]
    for inc in re.findall(r'include\("([^"]+)"\)', main):
        assert os.path.exists(os.path.join(SRC, inc)), inc


def test_every_ccall_matches_the_header():
    protos = header_prototypes()
    srcs = sources()
    for name, txt in srcs.items():
        if name != "lib.jl":
            assert "ccall((:pl_" not in txt, f"{name}: the C ABI is bound in src/lib.jl only"
    calls = ccalls(srcs["lib.jl"])
    assert len(calls) >= 25, "lib.jl binds the hot path"
    bound = set()
    for name, ret, args in calls:
        assert name in protos, f"{name} is ccall'ed by the package but not declared in include/plb200.h"
        cret, ctypes_ = protos[name]
        assert RET2JL[cret] == ret, f"{name}: return type {ret} vs C {cret}"
        want = [C2JL[t] for t in ctypes_]
        assert args == want, f"{name}: lib.jl passes {args}, header wants {want}"
        bound.add(name)
    for required in ("pl_init", "pl_init_all", "pl_finalize", "pl_gram_dot", "pl_gram_rbf", "pl_gram_cross", "pl_normal_eq", "pl_normal_eq_blocks",
                     "pl_neq_begin", "pl_neq_add", "pl_neq_finish", "pl_cov", "pl_global_feature", "pl_gram_resident", "pl_gram_fetch_block"):
        assert required in bound, required


# signatures the reference defines itself (file:line in /root/reference/src): a package-level method with one of these heads is an overwrite
REFERENCE_IDENTICAL = [
    r"learn!\(\s*lp::UnivariateLinearProblem,\s*α::Real\s*\)",                       # Learning/linear-learn.jl:11
    r"learn!\(\s*lp::CovariateLinearProblem,\s*α::Real\s*\)",                        # :37
    r"learn!\(\s*lp::UnivariateLinearProblem,\s*ss::SubsetSelector,\s*α::Real",      # :82
    r"learn!\(\s*lp::CovariateLinearProblem,\s*ss::SubsetSelector,\s*α::Real",       # :123
    r"learn!\(\s*lp::UnivariateLinearProblem,\s*ws::Vector,\s*int::Bool\s*\)",       # :178
    r"learn!\(\s*lp::CovariateLinearProblem,\s*ws::Vector,\s*int::Bool",             # :226
    r"ooc_learn!\(\s*lb::InteratomicPotentials\.LinearBasisPotential",               # :301
    r"compute_eigen\(\s*d::Vector\{T\}\s*\)\s*where",                                # DimensionReduction/dimension_reduction.jl:11
    r"fit!\(\s*ds::DataSet,\s*pca::PCAState\s*\)",                                   # DimensionReduction/pca_state.jl:20
]


def overrides_block(txt):
    a = txt.index("const OVERRIDES = quote")
    b = txt.rindex("\nend")
    return txt[:a], txt[a:b + 4], txt[b + 4:]


def test_no_reference_method_is_overwritten_at_package_level():
    srcs = sources()
    before, block, after = overrides_block(srcs["overrides.jl"])
    outside = "\n".join(v for k, v in srcs.items() if k != "overrides.jl") + before + after
    outside_code = "\n".join(l.split("#", 1)[0] for l in outside.splitlines())  # comments may quote signatures
    for pat in REFERENCE_IDENTICAL:
        assert not re.search(pat, outside_code), f"method overwriting at package level: {pat}"
        assert re.search(pat, block), f"the activation block must carry {pat}"
    # no other route to extending the reference's functions from package level
    assert not re.search(r"^\s*import\s+PotentialLearning", outside_code, flags=re.M), "`import PotentialLearning: f` + a method definition would be piracy at load time"
    assert not re.search(r"^\s*(function\s+)?PotentialLearning\.\w+!?\(.*\)\s*(=|$)", outside_code, flags=re.M)
    main = srcs["PotentialLearningB200.jl"]
    assert main.count("jl_generating_output") >= 2, "activate!() and __init__ are no-ops while precompiling"
    assert "Core.eval(PotentialLearning, OVERRIDES)" in main


def test_overrides_cover_the_reference_entry_points():
    _, block, _ = overrides_block(sources()["overrides.jl"])
    for sig in ("function KernelMatrix(F::Vector{Vector{Float64}}, k::Union{DotProduct,RBF,InverseMultiquadric})",
                "function KernelMatrix(F1::Vector{Vector{Float64}}, F2::Vector{Vector{Float64}}",
                "compute_features(ds::DataSet, f::GlobalMean", "function compute_divergence(x::Vector{Vector{Float64}}, div::KernelSteinDiscrepancy)"):
        assert sig in block, sig
    # the default eigensolver stays LAPACK (`eigen(Symmetric(Q))`, dimension_reduction.jl:13); the cuSOLVER path is opt-in
    dimred = sources()["dimred.jl"]
    assert 'get(ENV, "PLB200_EIGEN", "lapack") == "gpu" ? eigen_gpu(Q) : eigen(Symmetric(Q))' in dimred
    kernels = sources()["kernels.jl"]
    assert "Symmetric(K)" in kernels and "result_zeros(n, n)" in kernels
    learning = sources()["learning.jl"]
    assert "as_block(M::Matrix{Float64}) = M" in learning, "lp.dB[i] is passed as it lies in memory (no Matrix{Float64}(...) copy)"
