# LQCDB200.jl -- Julia-side shim that puts liblqcd_b200.so (include/lqcd_b200.h) behind the generic functions
# LatticeQCD.jl calls on its Dirac-solve hot path.  Thin by design: every method is a `ccall` into the C ABI.
#
# STATUS: written against the reference's call sites; NOT executed in the build image (no Julia runtime, and
# LatticeDiracOperators.jl / Gaugefields.jl sources are not vendored -- SURVEY.md section 0).  The Python twin
# (latticeqcd.jl_b200/lqcd_b200/) binds the same symbols with the same argument order and is what the tests
# and bench.py exercise.  Upstream type names marked [UPSTREAM-RECALL] must be checked against the installed
# package versions (LatticeDiracOperators 0.6.x, Gaugefields 0.4-0.7; /root/reference/Project.toml:24,27).
#
# Selection (SURVEY.md 8b, option 1 -- run_LQCD() stays unchanged): with ENV["LQCD_B200"] = "1" the method
#     Dirac_operator(U::Vector{<:AbstractGaugefields{3,4}}, x, params)           (src/system/universe.jl:137)
# returns a B200Dirac instead of upstream's CPU operator.  Link fields and pseudofermion fields remain the
# upstream CPU containers (the gauge sector keeps using them, src/md/AbstractMD.jl:78-118); links are mirrored
# to the device when the operator is built or re-bound (`D(U)`), fermion fields are copied in/out per call
# (a solve is hundreds of Dslash applications, the copies are noise).
module LQCDB200

using LinearAlgebra
import LinearAlgebra: mul!, dot
import Gaugefields: AbstractGaugefields                                   # [UPSTREAM-RECALL]
import LatticeDiracOperators                                               # [UPSTREAM-RECALL]
import LatticeDiracOperators: Dirac_operator, solve_DinvX!, FermiAction, calc_UdSfdU!, evaluate_FermiAction,
                              gauss_sampling_in_action!, sample_pseudofermions!, AbstractFermionfields

const LIB = get(ENV, "LQCD_B200_LIB", joinpath(@__DIR__, "..", "liblqcd_b200.so"))

const WILSON, STAGGERED = Cint(0), Cint(1)
const OP_D, OP_DDAG, OP_DDAGD = Cint(0), Cint(1), Cint(2)
const SOLVER_CG, SOLVER_CGNR, SOLVER_BICGSTAB = Cint(0), Cint(1), Cint(2)

# typedef struct { int kind; double kappa, r, mass, csw; double bc[4]; } lqcd_op;      (include/lqcd_b200.h)
struct LqcdOp
    kind::Cint
    kappa::Cdouble
    r::Cdouble
    mass::Cdouble
    csw::Cdouble
    bc::NTuple{4,Cdouble}
end

# ---- errors: the reference's convention is error(...) (universe.jl:73-75,130) ------------------------------
function check(ctx::Ptr{Cvoid}, st::Cint)
    st == 0 && return nothing
    msg = unsafe_string(ccall((:lqcd_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
    error("liblqcd_b200 error $st: $msg")          # LQCD_ERR_NOCONV (4) mirrors upstream's "The CG is not converged!"
end

# ---- context: one per process / GPU ---------------------------------------------------------------------------
mutable struct B200Context
    h::Ptr{Cvoid}
    dims::NTuple{4,Int}
    function B200Context(dims::NTuple{4,Int}; procgrid=(1, 1, 1, 1), rank=0, device=0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        d = Cint[dims...]; p = Cint[procgrid...]
        st = ccall((:lqcd_ctx_create, LIB), Cint, (Ptr{Cint}, Ptr{Cint}, Cint, Cint, Ref{Ptr{Cvoid}}), d, p, rank, device, out)
        check(C_NULL, st)
        ctx = new(out[], dims)
        finalizer(c -> ccall((:lqcd_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), c.h), ctx)
        return ctx
    end
end
const CONTEXTS = Dict{NTuple{4,Int},B200Context}()
context(dims) = get!(() -> B200Context(dims), CONTEXTS, dims)

# ---- device pseudofermion handle ------------------------------------------------------------------------------
mutable struct B200Field
    ctx::B200Context
    h::Ptr{Cvoid}
    kind::Cint
    function B200Field(ctx::B200Context, kind::Cint)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ctx.h, ccall((:lqcd_fermion_alloc, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}), ctx.h, kind, out))
        f = new(ctx, out[], kind)
        finalizer(x -> ccall((:lqcd_fermion_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), x.ctx.h, x.h), f)
        return f
    end
end

# Upstream CPU fields keep their data in `.f` (Wilson nowing: ComplexF64[NC,NX,NY,NZ,NT,4]; staggered with wing:
# [NC,NX+2,NY+2,NZ+2,NT+2,1]) [UPSTREAM-RECALL]; `wing(x)` is the halo width the library must strip / restore.
wing(x) = hasproperty(x, :NDW) ? Int(x.NDW) : 0
function upload!(d::B200Field, x::AbstractFermionfields)
    GC.@preserve x check(d.ctx.h, ccall((:lqcd_fermion_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{ComplexF64}, Cint),
                                        d.ctx.h, d.h, pointer(x.f), wing(x)))
end
function download!(x::AbstractFermionfields, d::B200Field)
    GC.@preserve x check(d.ctx.h, ccall((:lqcd_fermion_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{ComplexF64}, Cint),
                                        d.ctx.h, d.h, pointer(x.f), wing(x)))
end

# ---- the operator ---------------------------------------------------------------------------------------------
mutable struct B200Dirac{Mode}                 # Mode = :D, :Ddag, :DdagD
    ctx::B200Context
    op::LqcdOp
    eps::Float64
    maxsteps::Int
    method::Cint
    verbose::Int
    scratch::Vector{B200Field}                 # device twins of the host fields passed to mul!/solve
    cpu_template::Any                          # a host pseudofermion field (for similar())
    params::Dict
end

function upload_links!(ctx::B200Context, U::Vector{<:AbstractGaugefields{3,4}})
    # U[mu].U :: Array{ComplexF64,6} of size (NC,NC,NX+2w,NY+2w,NZ+2w,NT+2w) [UPSTREAM-RECALL: field name `U`, width `NDW`]
    w = hasproperty(U[1], :NDW) ? Int(U[1].NDW) : 0
    ptrs = [pointer(U[mu].U) for mu = 1:4]
    GC.@preserve U check(ctx.h, ccall((:lqcd_gauge_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{ComplexF64}}, Cint, Cint), ctx.h, ptrs, 3, w))
end

"Dirac_operator(U, x, params) -- src/system/universe.jl:103-137 builds exactly this params Dict."
function Dirac_operator(U::Vector{<:AbstractGaugefields{3,4}}, x, params::Dict)
    get(ENV, "LQCD_B200", "0") == "1" ||
        return invoke(Dirac_operator, Tuple{Array{<:AbstractGaugefields,1},Any,Any}, U, x, params)   # upstream CPU path
    name = params["Dirac_operator"]
    bc = Tuple(Float64.(get(params, "boundarycondition", [1, 1, 1, -1])))        # parameter_structs.jl:133
    op = if name == "Wilson"
        LqcdOp(WILSON, params["κ"], get(params, "r", 1.0), 0.0, 0.0, bc)
    elseif name == "staggered"
        LqcdOp(STAGGERED, 0.0, 1.0, params["mass"], 0.0, bc)
    else
        error("Dirac_operator = $name is not on the B200 path (Wilson, staggered)")
    end
    ctx = context((U[1].NX, U[1].NY, U[1].NZ, U[1].NT))
    upload_links!(ctx, U)
    method = Dict("bicg" => SOLVER_CGNR, "bicgstab" => SOLVER_BICGSTAB, "preconditiond_bicgstab" => SOLVER_BICGSTAB)[get(params, "method_CG", "bicg")]
    return B200Dirac{:D}(ctx, op, get(params, "eps_CG", 1e-19), get(params, "MaxCGstep", 3000), method,
                         get(params, "verbose_level", 1), [B200Field(ctx, op.kind) for _ = 1:4], x, params)
end

# D(U): re-bind (re-upload) the links -- measure_Pion_correlator.jl:338, and every MD step via calc_UdSfdU!
(D::B200Dirac)(U) = (upload_links!(D.ctx, U); D)
Base.adjoint(D::B200Dirac{:D}) = B200Dirac{:Ddag}(D.ctx, D.op, D.eps, D.maxsteps, D.method, D.verbose, D.scratch, D.cpu_template, D.params)
Base.adjoint(D::B200Dirac{:Ddag}) = B200Dirac{:D}(D.ctx, D.op, D.eps, D.maxsteps, D.method, D.verbose, D.scratch, D.cpu_template, D.params)
DdagD(D::B200Dirac{:D}) = B200Dirac{:DdagD}(D.ctx, D.op, D.eps, D.maxsteps, D.method, D.verbose, D.scratch, D.cpu_template, D.params)
LatticeDiracOperators.DdagD_operator(U, x, params) = DdagD(Dirac_operator(U, x, params))               # [UPSTREAM-RECALL]
mode(::B200Dirac{:D}) = OP_D; mode(::B200Dirac{:Ddag}) = OP_DDAG; mode(::B200Dirac{:DdagD}) = OP_DDAGD

"LinearAlgebra.mul!(y, D, x) -- measure_Pion_correlator.jl:379"
function mul!(y::AbstractFermionfields, D::B200Dirac, x::AbstractFermionfields)
    dx, dy = D.scratch[1], D.scratch[2]
    # one pipelined upload + Dslash + download (H2D, kernels and D2H overlap over slabs of t-slices); x.f / y.f are the
    # host arrays [UPSTREAM-RECALL: field name `f`], wing width as in upload!/download!
    w = hasproperty(x, :NDW) ? Int(x.NDW) : 0
    GC.@preserve x y check(D.ctx.h, ccall((:lqcd_dslash_host, LIB), Cint,
        (Ptr{Cvoid}, Ref{LqcdOp}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint, Cint),
        D.ctx.h, D.op, dy.h, dx.h, pointer(y.f), pointer(x.f), mode(D), w))
    return y
end

"solve_DinvX!(y, A, x): A y = x, y is the initial guess -- measure_Pion_correlator.jl:399, measure_chiral_condensate.jl:182"
function solve_DinvX!(y::AbstractFermionfields, A::B200Dirac, x::AbstractFermionfields)
    dx, dy = A.scratch[1], A.scratch[2]
    upload!(dx, x); upload!(dy, y)
    method = mode(A) == OP_DDAGD ? SOLVER_CG : A.method
    iters = Ref{Cint}(0); rs = Ref{Cdouble}(0.0)
    hist = A.verbose >= 3 ? zeros(Float64, A.maxsteps + 1) : Float64[]
    st = ccall((:lqcd_solve, LIB), Cint,
               (Ptr{Cvoid}, Ref{LqcdOp}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cdouble, Cint, Ref{Cint}, Ref{Cdouble}, Ptr{Cdouble}),
               A.ctx.h, A.op, dy.h, dx.h, method, mode(A), A.eps, A.maxsteps, iters, rs, isempty(hist) ? C_NULL : pointer(hist))
    check(A.ctx.h, st)
    A.verbose >= 3 && foreach(i -> println("$(i-1)-th eps: $(hist[i])"), 1:iters[]+1)     # upstream println_verbose_level3
    download!(y, dy)
    return y
end

"""
    solve_DinvX!(ys::Vector, A, xs::Vector)

All right-hand sides in lock step, every link fetched from HBM serving a group of them (lqcd_solve_multi, csrc/mrhs.cu): the loop
`map(i -> calc_quark_propagators_point_source_each(m, U, D, i, stvec), 1:NC*m.Nspinor)` of measure_Pion_correlator.jl:333-349 and
the `for ir = 1:Nr` loop of measure_chiral_condensate.jl:176-182 become ONE call.  Per right-hand side the result (and for the
default "bicg" method the iteration count) is that of its own `solve_DinvX!(y, A, x)`.
"""
function solve_DinvX!(ys::Vector{<:AbstractFermionfields}, A::B200Dirac, xs::Vector{<:AbstractFermionfields})
    n = length(ys); @assert n == length(xs) && n <= 16
    dx = [B200Field(A.ctx, A.op.kind) for _ = 1:n]; dy = [B200Field(A.ctx, A.op.kind) for _ = 1:n]
    for j = 1:n; upload!(dx[j], xs[j]); upload!(dy[j], ys[j]); end
    method = mode(A) == OP_DDAGD ? SOLVER_CG : A.method
    iters = zeros(Cint, n); rs = zeros(Cdouble, n)
    check(A.ctx.h, ccall((:lqcd_solve_multi, LIB), Cint,
        (Ptr{Cvoid}, Ref{LqcdOp}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Cint, Cint, Cint, Cdouble, Cint, Ptr{Cint}, Ptr{Cdouble}),
        A.ctx.h, A.op, [f.h for f in dy], [f.h for f in dx], n, method, mode(A), A.eps, A.maxsteps, iters, rs))
    for j = 1:n; download!(ys[j], dy[j]); end
    return ys
end
"mul!(ys, D, xs) for several fields in one pass over the links (lqcd_dslash_multi)"
function mul!(ys::Vector{<:AbstractFermionfields}, D::B200Dirac, xs::Vector{<:AbstractFermionfields})
    n = length(ys); @assert n == length(xs) && n <= 16
    dx = [B200Field(D.ctx, D.op.kind) for _ = 1:n]; dy = [B200Field(D.ctx, D.op.kind) for _ = 1:n]
    for j = 1:n; upload!(dx[j], xs[j]); end
    check(D.ctx.h, ccall((:lqcd_dslash_multi, LIB), Cint, (Ptr{Cvoid}, Ref{LqcdOp}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Cint, Cint),
                         D.ctx.h, D.op, [f.h for f in dy], [f.h for f in dx], n, mode(D)))
    for j = 1:n; download!(ys[j], dy[j]); end
    return ys
end

"""
    measure_chiral_condensate_b200(D, r_template; Nr=10, factor=1.0)

The Nr noise solves of measure(::Chiral_condensate_measurement) (measure_chiral_condensate.jl:164-204) as one batched solve:
pbp = real(sum_ir dot(r_ir, D^-1 r_ir) / Nr) / NV * factor with Z4 noise r_ir (the reference's Z4_distribution_fermi!, host RNG).
"""
function measure_chiral_condensate_b200(D::B200Dirac, r_template; Nr=10, factor=1.0)
    rs = [similar(r_template) for _ = 1:Nr]; ps = [similar(r_template) for _ = 1:Nr]
    for ir = 1:Nr
        LatticeDiracOperators.clear_fermion!(ps[ir])
        LatticeDiracOperators.Z4_distribution_fermi!(rs[ir])                                                    # [UPSTREAM-RECALL]
    end
    for j = 1:16:Nr
        k = min(j + 15, Nr)
        solve_DinvX!(ps[j:k], D, rs[j:k])
    end
    NV = prod(D.ctx.dims)
    return real(sum(dot(rs[ir], ps[ir]) for ir = 1:Nr) / Nr) / NV * factor
end

# ---- fermion action (Wilson two-flavour / staggered Nf=8 form; RHMC fractions go through lqcd_multishift_cg) ----
struct B200FermiAction
    D::B200Dirac{:D}
    _temporary_fermionfields::Vector{Any}        # src/md/standardMD.jl:50 does similar(fermi_action._temporary_fermionfields[1])
    parameters_action::Dict
end
"FermiAction(D, parameters_action) -- src/system/universe.jl:138; staggered Nf not in {4, 8} -> rational HMC like upstream (README.md:132)"
function FermiAction(D::B200Dirac{:D}, parameters_action)
    temps = [similar(D.cpu_template) for _ = 1:4]
    Nf = get(parameters_action, "Nf", D.op.kind == STAGGERED ? 8 : 2)
    (D.op.kind == STAGGERED && !(Nf in (4, 8))) || return B200FermiAction(D, temps, parameters_action)
    # partial fractions x^p ~ c0 + sum_j c[j] / (x + s[j]) on the spectrum of D^dag D, [m^2, m^2 + 16]: upstream takes them from
    # AlgRemez_jll; any (c0, c, s) triple of that form works -- passed in as parameters_action["rational_action"] (p = -Nf/8)
    # and ["rational_heatbath"] (p = +Nf/16), e.g. computed once with lqcd_b200/rhmc.py:rational_approx.
    haskey(parameters_action, "rational_action") && haskey(parameters_action, "rational_heatbath") ||
        error("B200 RHMC: parameters_action needs \"rational_action\" and \"rational_heatbath\" = (c0, c::Vector, s::Vector)")
    a0, a, b = parameters_action["rational_action"]; h0, h, hb = parameters_action["rational_heatbath"]
    return B200RHMCAction(D, Float64(a0), Float64.(a), Float64.(b), Float64(h0), Float64.(h), Float64.(hb), temps)
end

# staggered Nf = 4: even-site pseudofermions (odd sites zeroed after the Gaussian sampling and after D^dag) [UPSTREAM-RECALL,
# SURVEY.md App. C.7]; Nf = 8 and Wilson: all sites.  Other staggered Nf -> B200RHMCAction below.
even_only(fa::B200FermiAction) = fa.D.op.kind == STAGGERED && get(fa.parameters_action, "Nf", 8) == 4
"gauss_sampling_in_action!(xi, U, fa) -- src/md/standardMD.jl:95 (host RNG stays the reference's, seeded by Random.seed!, lqcd.jl:61)"
function gauss_sampling_in_action!(ξ, U, fa::B200FermiAction)
    LatticeDiracOperators.gauss_distribution_fermion!(ξ)                                                        # [UPSTREAM-RECALL]
    if even_only(fa)
        # eta = P_even D^dag xi0 has the right heat-bath covariance (D^dag D)_ee only for xi0 on ALL sites, but is then a projection:
        # hand back xi = D (D^dag D)^-1 eta, for which P_even D^dag xi = eta and dot(xi, xi) (Sfold, standardHMC.jl:54) equals
        # eta^dag (D^dag D)^-1 eta, so the reference's update! stays exact (see lqcd_b200/api.py FermiActionB200)
        D = fa.D(U)
        η0, X = fa._temporary_fermionfields[3], fa._temporary_fermionfields[1]
        mul!(η0, adjoint(D), ξ); LatticeDiracOperators.clear_fermion!(η0, false)                                # [UPSTREAM-RECALL] evensite = false
        LatticeDiracOperators.clear_fermion!(X)
        solve_DinvX!(X, DdagD(D), η0)
        mul!(ξ, D, X)
    end
    return ξ
end
"sample_pseudofermions!(eta, U, fa, xi): eta = D^dag xi -- src/md/standardMD.jl:96"
function sample_pseudofermions!(η, U, fa::B200FermiAction, ξ)
    mul!(η, adjoint(fa.D(U)), ξ)
    even_only(fa) && LatticeDiracOperators.clear_fermion!(η, false)
    return η
end
# even-site action: its (D^dag D) solves run on checkerboarded half fields (lqcd_solve_staggered_even) while this switch is on
function with_even_site_solves(f, fa::B200FermiAction)
    on = even_only(fa) && get(fa.parameters_action, "half_field_solver", true)
    on && check(fa.D.ctx.h, ccall((:lqcd_set_staggered_even_solve, LIB), Cint, (Ptr{Cvoid}, Cint), fa.D.ctx.h, 1))
    try
        return f()
    finally
        on && check(fa.D.ctx.h, ccall((:lqcd_set_staggered_even_solve, LIB), Cint, (Ptr{Cvoid}, Cint), fa.D.ctx.h, 0))
    end
end
"evaluate_FermiAction(fa, U, eta) = eta^dag (D^dag D)^-1 eta -- src/updates/standardHMC.jl:69-71"
function evaluate_FermiAction(fa::B200FermiAction, U, η)
    X = fa._temporary_fermionfields[1]
    LatticeDiracOperators.clear_fermion!(X)
    with_even_site_solves(fa) do
        solve_DinvX!(X, DdagD(fa.D(U)), η)
    end
    return real(dot(η, X))
end

"calc_UdSfdU!(UdSfdU, fa, U, eta) -- src/md/AbstractMD.jl:129; CG, Y = D X and the outer products all run on the device"
function calc_UdSfdU!(UdSfdU::Vector{<:AbstractGaugefields{3,4}}, fa::B200FermiAction, U, η)
    D = fa.D(U)
    dη = D.scratch[1]
    upload!(dη, η)
    outs = [pointer(UdSfdU[mu].U) for mu = 1:4]                 # temporaries from get_temp(temps, Dim), AbstractMD.jl:123
    w = hasproperty(UdSfdU[1], :NDW) ? Int(UdSfdU[1].NDW) : 0
    iters = Ref{Cint}(0); act = Ref{Cdouble}(0.0)
    with_even_site_solves(fa) do
        GC.@preserve UdSfdU check(D.ctx.h, ccall((:lqcd_fermion_force, LIB), Cint,
            (Ptr{Cvoid}, Ref{LqcdOp}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Ptr{Ptr{ComplexF64}}, Cint, Ref{Cint}, Ref{Cdouble}),
            D.ctx.h, D.op, dη.h, C_NULL, D.eps, D.maxsteps, outs, w, iters, act))
    end
    return nothing
end

# ---- rational HMC (staggered Nf not in {4, 8}; README.md:132, test/test_Nf2.toml) -----------------------------------------
# x^(-Nf/8) ~ a0 + sum_j a[j]/(x + b[j]) (coefficients from AlgRemez_jll as upstream, or any partial-fraction fit): the MD
# force sum_j a[j] * force(X_j, Y_j) is accumulated on the device, X_j from ONE lqcd_multishift_cg, Y_j = D X_j.
struct B200RHMCAction
    D::B200Dirac{:D}
    a0::Float64; a::Vector{Float64}; b::Vector{Float64}          # action approximation x^(-Nf/8), b ascending
    h0::Float64; h::Vector{Float64}; hb::Vector{Float64}         # heat-bath approximation x^(+Nf/16)
    _temporary_fermionfields::Vector{Any}
end

"y = c0 x + sum_j c[j] (D^dag D + s[j])^-1 x with ONE multi-shift CG on the device (lqcd_rational_apply); returns Re<x, y>"
function rational_apply!(y, D::B200Dirac, x, c0, c, s)
    dx, dy = D.scratch[1], D.scratch[2]
    upload!(dx, x)
    iters = Ref{Cint}(0); d = Ref{Cdouble}(0.0)
    check(D.ctx.h, ccall((:lqcd_rational_apply, LIB), Cint,
        (Ptr{Cvoid}, Ref{LqcdOp}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cdouble, Cint, Ref{Cint}, Ref{Cdouble}),
        D.ctx.h, D.op, dy.h, dx.h, c0, c, s, length(c), D.eps, D.maxsteps, iters, d))
    download!(y, dy)
    return d[]
end
"gauss_sampling_in_action!(xi, U, fa) -- src/md/standardMD.jl:95"
gauss_sampling_in_action!(ξ, U, fa::B200RHMCAction) = LatticeDiracOperators.gauss_distribution_fermion!(ξ)     # [UPSTREAM-RECALL]
"sample_pseudofermions!(eta, U, fa, xi): eta = (D^dag D)^{Nf/16} xi -- src/md/standardMD.jl:96"
function sample_pseudofermions!(η, U, fa::B200RHMCAction, ξ)
    rational_apply!(η, fa.D(U), ξ, fa.h0, fa.h, fa.hb)
    return η
end

function calc_UdSfdU!(UdSfdU::Vector{<:AbstractGaugefields{3,4}}, fa::B200RHMCAction, U, η)
    D = fa.D(U)
    dη = D.scratch[1]; upload!(dη, η)
    outs = [pointer(UdSfdU[mu].U) for mu = 1:4]
    w = hasproperty(UdSfdU[1], :NDW) ? Int(UdSfdU[1].NDW) : 0
    iters = Ref{Cint}(0)
    GC.@preserve UdSfdU check(D.ctx.h, ccall((:lqcd_fermion_force_rational, LIB), Cint,
        (Ptr{Cvoid}, Ref{LqcdOp}, Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cdouble, Cint, Ptr{Ptr{ComplexF64}}, Cint, Ref{Cint}),
        D.ctx.h, D.op, dη.h, fa.a, fa.b, length(fa.b), D.eps, D.maxsteps, outs, w, iters))
    return nothing
end
"evaluate_FermiAction(fa, U, eta) = eta^dag r(D^dag D) eta, r ~ x^(-Nf/8): one multi-shift CG on the device"
evaluate_FermiAction(fa::B200RHMCAction, U, η) = rational_apply!(fa._temporary_fermionfields[1], fa.D(U), η, fa.a0, fa.a, fa.b)

# ---- gauge configurations in the reference's file formats straight to / from the device links (csrc/gauge_io.cu) -----------------
# `initial = "<file>"` + loadU_format (universe.jl:62-68) and saveU_format (lqcd.jl:236-242).  The host-array forms
# (load_BridgeText!, ILDG + load_gaugefield!, save_binarydata, save_textdata) stay Gaugefields.jl's; these skip the host arrays.
const IO_FORMATS = Dict("ILDG" => 0, "BridgeText" => 1)
load_gaugefield_device!(ctx::B200Context, filename::String, loadU_format::String="ILDG") =
    check(ctx.h, ccall((:lqcd_gauge_load, LIB), Cint, (Ptr{Cvoid}, Cstring, Cint), ctx.h, filename, IO_FORMATS[loadU_format]))
save_gaugefield_device(ctx::B200Context, filename::String, saveU_format::String="ILDG") =
    check(ctx.h, ccall((:lqcd_gauge_save, LIB), Cint, (Ptr{Cvoid}, Cstring, Cint), ctx.h, filename, IO_FORMATS[saveU_format]))

# ---- device-resident molecular dynamics (src/md/standardMD.jl:103-165, src/md/AbstractMD.jl:78-135) -------------------------
# runMD!(U, md) for a StandardMD whose fermi_action is a B200FermiAction (or quenched): the links are uploaded once, momenta
# are sampled on the device, U_update! / P_update! / P_update_fermion! run as kernels (lqcd_md_trajectory), and U is
# downloaded at the end.  md.p (host momenta) is bypassed: Sp_old / Sp_new come from lqcd_md_kinetic.
function runMD_b200!(U::Vector{<:AbstractGaugefields{3,4}}, ctx::B200Context, β, Δτ, MDsteps; fa::Union{Nothing,B200FermiAction,B200RHMCAction}=nothing,
                     η=nothing, SextonWeingargten=false, Nsw=2, seed=rand(UInt64))
    upload_links!(ctx, U)
    check(ctx.h, ccall((:lqcd_md_momenta_gaussian, LIB), Cint, (Ptr{Cvoid}, UInt64), ctx.h, seed))
    K0 = Ref{Cdouble}(0.0); check(ctx.h, ccall((:lqcd_md_kinetic, LIB), Cint, (Ptr{Cvoid}, Ref{Cdouble}), ctx.h, K0))
    its = Ref{Clonglong}(0)
    if fa === nothing
        check(ctx.h, ccall((:lqcd_md_trajectory, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cint, Cint, Cdouble, Cint, Ref{Clonglong}),
            ctx.h, C_NULL, C_NULL, β, Δτ, MDsteps, SextonWeingargten ? Nsw : 0, 0.0, 1, its))
    elseif fa isa B200RHMCAction                      # every fermion force = one multi-shift CG + accumulated outer products
        dη = fa.D.scratch[1]; upload!(dη, η)
        check(ctx.h, ccall((:lqcd_md_trajectory_rational, LIB), Cint,
            (Ptr{Cvoid}, Ref{LqcdOp}, Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cdouble, Cdouble, Cint, Cint, Cdouble, Cint, Ref{Clonglong}),
            ctx.h, fa.D.op, dη.h, fa.a, fa.b, length(fa.b), β, Δτ, MDsteps, SextonWeingargten ? Nsw : 0, fa.D.eps, fa.D.maxsteps, its))
    else
        dη = fa.D.scratch[1]; upload!(dη, η)
        check(ctx.h, ccall((:lqcd_md_trajectory, LIB), Cint,
            (Ptr{Cvoid}, Ref{LqcdOp}, Ptr{Cvoid}, Cdouble, Cdouble, Cint, Cint, Cdouble, Cint, Ref{Clonglong}),
            ctx.h, fa.D.op, dη.h, β, Δτ, MDsteps, SextonWeingargten ? Nsw : 0, fa.D.eps, fa.D.maxsteps, its))
    end
    K1 = Ref{Cdouble}(0.0); check(ctx.h, ccall((:lqcd_md_kinetic, LIB), Cint, (Ptr{Cvoid}, Ref{Cdouble}), ctx.h, K1))
    w = hasproperty(U[1], :NDW) ? Int(U[1].NDW) : 0
    ptrs = [pointer(U[mu].U) for mu = 1:4]
    GC.@preserve U check(ctx.h, ccall((:lqcd_gauge_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{ComplexF64}}, Cint, Cint), ctx.h, ptrs, 3, w))
    return (Sp_old = K0[], Sp_new = K1[], cg_iters = its[])
end

end # module
