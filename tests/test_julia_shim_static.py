"""Static consistency check of the Julia shim (julia/PotentialLearningB200.jl) against include/plb200.h: Julia is not installed in
this image, so the shim cannot be executed — but every `ccall` in it must name an exported function and pass exactly the argument
types the header declares, in order. Catches drift between the C ABI and the reference-side binding."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "plb200.h")
SHIM = os.path.join(ROOT, "potentiallearning.jl_b200", "julia", "PotentialLearningB200.jl")

C2JL = {
    "const double*": "Ptr{Cdouble}", "double*": "Ptr{Cdouble}", "const int64_t*": "Ptr{Int64}", "int64_t*": "Ptr{Int64}",
    "const double* const*": "Ptr{Ptr{Cdouble}}", "int64_t": "Int64", "int": "Cint", "double": "Cdouble", "void*": "Ptr{Cvoid}",
    "const int*": "Ptr{Cint}",
}
RET2JL = {"int": "Cint", "void": "Cvoid", "const char*": "Cstring", "int64_t": "Int64"}


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


def shim_ccalls():
    src = open(SHIM).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(pl_\w+),\s*libplb200\),\s*(\w+),\s*\(([^)]*)\)", src, flags=re.S):
        args = [a.strip() for a in m.group(3).replace("\n", " ").split(",") if a.strip()]
        calls.append((m.group(1), m.group(2), args))
    return calls


def test_every_ccall_matches_the_header():
    protos = header_prototypes()
    calls = shim_ccalls()
    assert len(calls) >= 12, "the shim binds the hot path"
    for name, ret, args in calls:
        assert name in protos, f"{name} is ccall'ed by the shim but not declared in include/plb200.h"
        cret, ctypes_ = protos[name]
        assert RET2JL[cret] == ret, f"{name}: return type {ret} vs C {cret}"
        want = [C2JL[t] for t in ctypes_]
        assert args == want, f"{name}: shim passes {args}, header wants {want}"


def test_shim_covers_the_reference_entry_points():
    src = open(SHIM).read()
    for sig in ("function KernelMatrix(F::Vector{Vector{Float64}}, k::DotProduct)", "function KernelMatrix(F::Vector{Vector{Float64}}, k::RBF)",
                "function KernelMatrix(F1::Vector{Vector{Float64}}, F2::Vector{Vector{Float64}}", "function learn!(lp::UnivariateLinearProblem, ws::Vector, int::Bool)",
                "function learn!(lp::CovariateLinearProblem, ws::Vector, int::Bool; λ::Real = 0.0)", "function learn!(lp::UnivariateLinearProblem, α::Real)",
                "compute_eigen(d::Vector{T}) where {T<:Vector{<:Real}}", "function fit!(ds::DataSet, pca::PCAState)"):
        assert sig in src, sig
    assert "Symmetric(K)" in src and "PL_FILL_JULIA_U" in src
