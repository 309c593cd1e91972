# CedarSimB200Ext.jl -- Julia glue between CedarSim's sweep API and libcedarb200.so (include/cedarb200.h).
#
# STATUS: written against the reference's API (src/sweeps.jl:390-502) and this repository's C ABI, but UNTESTED: the image
# this repository is built in has no Julia toolchain.  What IS tested here: the same entry points through ctypes
# (cedarsim.jl_b200/engine.py), the C harness tests/abi_smoke.c, and -- for this file specifically -- that the struct
# mirrors below have exactly the C layout (tests/test_abi.py parses this file and compares field offsets with
# `offsetof` from the C compiler).  Keep the `struct ... end` blocks in the plain form the test understands.
#
# What it overrides (reference file:line -> C entry point):
#   dc!(cs::CircuitSweep; kwargs...)          src/sweeps.jl:448        -> cb_dc
#   tran!(cs::CircuitSweep, tspan; kwargs...) src/sweeps.jl:457        -> cb_tran   (tspan explicit: the reference's method
#                                                                         has no tspan and its broadcast override is broken,
#                                                                         SURVEY.md 3.2)
#   CircuitSweep(circuit, iterator) compile   src/sweeps.jl:414-417    -> `python -m cedarsim.jl_b200.flatten` (front end,
#                                                                         once per sweep) + cb_circuit_load + cb_circuit_compile
#
# Usage:
#   using CedarSim, CedarSimB200Ext
#   cs   = B200Sweep("dff.cir", TandemSweep(...); outputs = ["q", "d"])     # netlist path + any SweepLike
#   sols = tran!(cs, (0.0, 6e-7); saveat = range(0, 6e-7; length = 1801), reltol = 1e-4)
#   sols[i](1.5e-7; idxs = "q"); sols[i].retcode
module CedarSimB200Ext

using CedarSim
import CedarSim: dc!, tran!, sweepvars

const lib = get(ENV, "CEDARB200_LIB", "libcedarb200")
const python = get(ENV, "CEDARB200_PYTHON", "python")

# ---- mirrors of include/cedarb200.h (layout asserted by tests/test_abi.py and by cb_options_init at run time) ----
struct cb_pref
    value::Cdouble
    col::Int32
    _pad::Int32
end

mutable struct cb_options
    struct_size::UInt32
    abi_version::UInt32
    temp::cb_pref
    gmin::cb_pref
    reltol::Cdouble
    vabstol::Cdouble
    iabstol::Cdouble
    nr_reltol::Cdouble
    nr_vabstol::Cdouble
    nr_iabstol::Cdouble
    dc_abstol::Cdouble
    dv_max::Cdouble
    max_newton_dc::Int32
    max_newton_tran::Int32
    method::Int32
    fixed_step::Int32
    dt::Cdouble
    dt_min::Cdouble
    dt_max::Cdouble
    gmin_steps::Int32
    skip_dc::Int32
    nr_rate_test::Int32
    value_rounds::Int32
    mixed_rounds::Int32
    source_steps::Int32
    t0_reinit::Int32
    pivot_repair::Int32
    pivot_growth_max::Cdouble
    cb_options() = new()
end

mutable struct cb_stats
    newton_iters::Int64
    lu_factors::Int64
    steps_accepted::Int64
    steps_rejected::Int64
    rounds::Int64
    kernel_launches::Int64
    solve_seconds::Cdouble
    h2d_seconds::Cdouble
    d2h_seconds::Cdouble
    eval_seconds::Cdouble
    newton_seconds::Cdouble
    value_rounds::Int64
    full_iters::Int64
    evalv_seconds::Cdouble
    newtonv_seconds::Cdouble
    pivot_fallbacks::Int64
    dc_source_stepped::Int64
    cb_stats() = new()
end

check(rc) = rc == 0 ? nothing : error("cedarb200 error $rc: " * unsafe_string(ccall((:cb_last_error, lib), Cstring, ())))

"Options with the library's defaults; fails (instead of corrupting memory) if this file's mirror is out of date."
function default_options(; kwargs...)
    o = cb_options()
    check(ccall((:cb_options_init, lib), Cint, (Ref{cb_options}, Csize_t), o, sizeof(cb_options)))
    for (k, v) in kwargs
        if k === :temp || k === :gmin
            setfield!(o, k, cb_pref(Float64(v), Int32(-1), Int32(0)))
        else
            setfield!(o, k, convert(fieldtype(cb_options, k), v))
        end
    end
    return o
end

const RETCODES = (:Success, :MaxIters, :InitialFailure, :DtLessThanMin, :Unstable)   # CB_ST_* -> SciMLBase.ReturnCode names

# ---- the sweep object: what CircuitSweep is for the reference, with the engine handles attached -------------------------
mutable struct B200Sweep
    iterator::Any                 # any CedarSim SweepLike
    circuit::Ptr{Cvoid}
    plan::Ptr{Cvoid}
    params::Matrix{Float64}       # (B, P): C layout [P][B]
    outputs::Vector{String}
    meta::Dict{String,Any}
    compile_seconds::Float64
end

Base.length(cs::B200Sweep) = length(cs.iterator)
Base.size(cs::B200Sweep) = size(cs.iterator)
sweepvars(cs::B200Sweep) = sweepvars(cs.iterator)

"Sweep points as a CSV the front end reads: header = swept names, one row per point in `collect(iterator)` order
(first axis fastest, src/sweeps.jl:261-268); `nothing` (SerialSweep's inactive variables) is not supported here."
function write_points(path, iterator)
    names = sort!(collect(sweepvars(iterator)))
    open(path, "w") do io
        println(io, join(string.(names), ","))
        for point in iterator
            d = Dict(point)
            println(io, join((repr(Float64(d[n])) for n in names), ","))
        end
    end
    return names
end

"""
    B200Sweep(netlist_path, iterator; outputs, device = 0, lang = "spice", cache_dir = nothing)

Compile once for the set of swept names (the reference does the same from `first(iterator)`, src/sweeps.jl:414-417):
run the front end on the deck, load the flat circuit + generated CUDA C, NVRTC-compile for sm_100a, create the plan and
bind every point's parameters.
"""
function B200Sweep(netlist_path::AbstractString, iterator; outputs::Vector{String}, device::Integer = 0, lang = "spice",
                   cache_dir = nothing)
    prefix = tempname()
    write_points(prefix * ".csv", iterator)
    run(`$python -m cedarsim.jl_b200.flatten $netlist_path --sweep $(prefix * ".csv") --outputs $(join(outputs, ",")) --out $prefix --lang $lang`)
    meta = _read_json(prefix * ".json")
    B, P = Int(meta["B"]), Int(meta["P"])
    params = Matrix{Float64}(undef, B, P)
    P > 0 && read!(prefix * ".params.f64", params)
    c = Ref{Ptr{Cvoid}}(C_NULL); p = Ref{Ptr{Cvoid}}(C_NULL); secs = Ref{Cdouble}(0.0)
    check(ccall((:cb_circuit_load, lib), Cint, (Cstring, Ref{Ptr{Cvoid}}), prefix * ".flatckt", c))
    check(ccall((:cb_circuit_compile, lib), Cint, (Ptr{Cvoid}, Cstring, Ref{Cdouble}), c[],
                cache_dir === nothing ? C_NULL : cache_dir, secs))
    check(ccall((:cb_plan_create, lib), Cint, (Ptr{Cvoid}, Int64, Cint, Ref{Ptr{Cvoid}}), c[], B, device, p))
    cs = B200Sweep(iterator, c[], p[], params, String.(meta["outputs"]), meta, secs[])
    GC.@preserve params check(ccall((:cb_plan_set_params, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), p[], params))
    finalizer(cs) do x
        ccall((:cb_plan_destroy, lib), Cvoid, (Ptr{Cvoid},), x.plan)
        ccall((:cb_circuit_destroy, lib), Cvoid, (Ptr{Cvoid},), x.circuit)
    end
    return cs
end

"""
    B200Sweep(spice_text, iterator, Val(:native); outputs, device = 0, cache_dir = nothing, base_dir = nothing)

The same with the LIBRARY's own netlist front end (`cb_netlist_flatten`, include/cedarb200.h): SPICE text and the sweep's
values go straight across the C ABI -- no Python process, no files.  For decks of R C L V I E G and subcircuits; decks
with MOSFETs / Verilog-A / behavioural sources are refused by the library with a message and take the method above.
`nothing` entries of a SerialSweep are passed as NaN (= keep the default, src/sweeps.jl:18-21).
"""
function B200Sweep(spice_text::AbstractString, iterator, ::Val{:native}; outputs::Vector{String}, device::Integer = 0,
                   cache_dir = nothing, base_dir = nothing)
    names = sort!(collect(string.(sweepvars(iterator))))
    points = collect(iterator)
    B = length(points)
    vals = Matrix{Float64}(undef, B, length(names))          # (B, n_sweep): C layout [n_sweep][B]
    for (i, point) in enumerate(points)
        d = Dict(string(k) => v for (k, v) in pairs(point))
        for (j, n) in enumerate(names)
            vals[i, j] = d[n] === nothing ? NaN : Float64(d[n])
        end
    end
    nl = Ref{Ptr{Cvoid}}(C_NULL); c = Ref{Ptr{Cvoid}}(C_NULL); p = Ref{Ptr{Cvoid}}(C_NULL); secs = Ref{Cdouble}(0.0)
    GC.@preserve vals check(ccall((:cb_netlist_flatten, lib), Cint,
        (Cstring, Cstring, Ptr{Cstring}, Cint, Ptr{Cdouble}, Int64, Ptr{Cstring}, Cint, Ref{Ptr{Cvoid}}),
        spice_text, base_dir === nothing ? C_NULL : base_dir, names, length(names), vals, B, outputs, length(outputs), nl))
    check(ccall((:cb_netlist_circuit, lib), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), nl[], c))
    check(ccall((:cb_circuit_compile, lib), Cint, (Ptr{Cvoid}, Cstring, Ref{Cdouble}), c[],
                cache_dir === nothing ? C_NULL : cache_dir, secs))
    P = Int(ccall((:cb_netlist_n_params, lib), Int32, (Ptr{Cvoid},), nl[]))
    params = Matrix{Float64}(undef, B, P)
    P > 0 && unsafe_copyto!(pointer(params), ccall((:cb_netlist_params, lib), Ptr{Cdouble}, (Ptr{Cvoid},), nl[]), B * P)
    ccall((:cb_netlist_destroy, lib), Cvoid, (Ptr{Cvoid},), nl[])
    check(ccall((:cb_plan_create, lib), Cint, (Ptr{Cvoid}, Int64, Cint, Ref{Ptr{Cvoid}}), c[], B, device, p))
    cs = B200Sweep(iterator, c[], p[], params, lowercase.(outputs), Dict{String,Any}("B" => B, "P" => P), secs[])
    GC.@preserve params check(ccall((:cb_plan_set_params, lib), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), p[], params))
    finalizer(cs) do x
        ccall((:cb_plan_destroy, lib), Cvoid, (Ptr{Cvoid},), x.plan)
        ccall((:cb_circuit_destroy, lib), Cvoid, (Ptr{Cvoid},), x.circuit)
    end
    return cs
end

# minimal JSON reader for the flat {"key": number | string | [..] | {..} | null} files the front end writes
function _read_json(path)
    s = read(path, String); i = Ref(1)
    ws() = (while i[] <= lastindex(s) && isspace(s[i[]]); i[] += 1; end)
    function val()
        ws(); ch = s[i[]]
        if ch == '{'
            d = Dict{String,Any}(); i[] += 1; ws()
            if s[i[]] == '}'; i[] += 1; return d; end
            while true
                k = val(); ws(); i[] += 1               # ':'
                d[k] = val(); ws()
                s[i[]] == ',' ? (i[] += 1) : (i[] += 1; return d)
            end
        elseif ch == '['
            a = Any[]; i[] += 1; ws()
            if s[i[]] == ']'; i[] += 1; return a; end
            while true
                push!(a, val()); ws()
                s[i[]] == ',' ? (i[] += 1) : (i[] += 1; return a)
            end
        elseif ch == '"'
            j = findnext('"', s, i[] + 1); str = s[i[]+1:j-1]; i[] = j + 1; return str
        elseif startswith(SubString(s, i[]), "null")
            i[] += 4; return nothing
        else
            j = i[]
            while j <= lastindex(s) && !(s[j] in (',', ']', '}', ' ', '\n')); j += 1; end
            x = parse(Float64, s[i[]:j-1]); i[] = j; return x
        end
    end
    return val()
end

# ---- results: what the reference's tests touch on a solution element (test/sweep.jl:336-339, test/gf180_dff.jl:28-33) --
struct B200DCSolution
    x::Vector{Float64}            # outputs of this point
    retcode::Symbol
    names::Vector{String}
end
Base.getindex(s::B200DCSolution, name::AbstractString) = s.x[findfirst(==(lowercase(name)), s.names)]

struct B200TranSolution
    t::Vector{Float64}
    y::Matrix{Float64}            # (S, O)
    retcode::Symbol
    names::Vector{String}
end
Base.getindex(s::B200TranSolution, name::AbstractString) = s.y[:, findfirst(==(lowercase(name)), s.names)]
function (s::B200TranSolution)(t::Real; idxs::AbstractString)      # sol(t; idxs = ...): linear interpolation on saveat
    k = clamp(searchsortedlast(s.t, t), 1, length(s.t) - 1)
    w = (t - s.t[k]) / (s.t[k+1] - s.t[k])
    col = findfirst(==(lowercase(idxs)), s.names)
    return (1 - w) * s.y[k, col] + w * s.y[k+1, col]
end

"dc!(cs): operating point of every sweep point, Array of size(cs) (reference: dc!.(cs.sys, cs), src/sweeps.jl:448)"
function dc!(cs::B200Sweep; abstol = 1e-10, kwargs...)
    B, O = length(cs), length(cs.outputs)
    opts = default_options(; dc_abstol = min(abstol, 1e-10), kwargs...)      # CedarDCOp: min(abstol, 1e-10), src/dcop.jl:101
    x = Matrix{Float64}(undef, B, O); status = Vector{Int32}(undef, B); st = cb_stats()
    GC.@preserve x status check(ccall((:cb_dc, lib), Cint,
        (Ptr{Cvoid}, Ref{cb_options}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}, Ref{cb_stats}), cs.plan, opts, x, C_NULL, status, st))
    return reshape([B200DCSolution(x[i, :], RETCODES[status[i]+1], cs.outputs) for i in 1:B], size(cs))
end

"tran!(cs, tspan; saveat): transient of every sweep point (reference: tran!.(cs.sys, cs), src/sweeps.jl:457)"
function tran!(cs::B200Sweep, tspan; saveat, reltol = 1e-3, abstol = 1e-6, kwargs...)
    B, O = length(cs), length(cs.outputs)
    ts = collect(Float64, saveat); S = length(ts)
    opts = default_options(; reltol = reltol, vabstol = abstol, kwargs...)
    y = Array{Float64,3}(undef, B, S, O); status = Vector{Int32}(undef, B); st = cb_stats()
    GC.@preserve y status ts check(ccall((:cb_tran, lib), Cint,
        (Ptr{Cvoid}, Cdouble, Cdouble, Ptr{Cdouble}, Int64, Ref{cb_options}, Ptr{Cdouble}, Ptr{Int32}, Ref{cb_stats}),
        cs.plan, Float64(tspan[1]), Float64(tspan[2]), ts, S, opts, y, status, st))
    return reshape([B200TranSolution(ts, y[i, :, :], RETCODES[status[i]+1], cs.outputs) for i in 1:B], size(cs))
end

export B200Sweep, default_options
end # module
