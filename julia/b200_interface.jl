#=
b200_interface.jl

`ccall` glue between SolarOscFitter's Julia host code and libsolarosc_b200.so, patterned on the
reference's NuFast glue (src/oscillations/nuFast_interface.jl:1-129): one module, a handle created
once after setup.jl has built its tables, thin wrappers that keep the call signatures of the
functions they stand in for.

NOT EXECUTED IN THIS REPOSITORY: the build image has no Julia.  The C ABI it binds is exercised
from Python (ctypes) by tests/; this file is the binding a SolarOscFitter maintainer drops into
src/oscillations/ (INTEGRATION.md has the step-by-step wiring and the Makefile rule).

Layout note: Julia arrays are column-major, so the handle is created with `col_major = 1` and every
matrix (response matrices R[Etrue, Ereco], night data/rates [cosz, Ereco], P_night [cosz, Etrue])
is passed and returned as it lies in Julia memory.  Batched outputs get the batch index LAST, e.g.
`ES_night[z, j, b]`.
=#
module B200Osc

using Base: @ccall
import ForwardDiff
import DensityInterface

const libb200 = joinpath(@__DIR__, "libsolarosc_b200.so")
const Handle = Ptr{Cvoid}
const handle = Ref{Handle}(C_NULL)

const ABI_VERSION = Int32(3)
const N_FIXED = 8

# ---- C structs (include/solarosc_b200.h) ---------------------------------------------------------
struct CDetectorExt                      # solarosc_detector_ext: inclusive / semi-inclusive / angular branches
    inclusive::Int32; angular::Int32; n_cos::Int32; reserved::Int32
    cc_f_above::Cdouble; cc_f_below::Cdouble
    resp_cc_incl::Ptr{Cdouble}; eff_cc_incl::Ptr{Cdouble}
    resp_cc_misid_nue::Ptr{Cdouble}; resp_cc_misid_nuother::Ptr{Cdouble}
    resp_es_angular::Ptr{Cdouble}; resp_bg_angular::Ptr{Cdouble}; angular_mask::Ptr{Cdouble}
end

struct CDetector
    es_mode::Int32; cc_mode::Int32
    n_ereco_es::Int32; n_ereco_cc::Int32
    index_es::Int32; index_cc::Int32
    unosc_es_nue_8b::Ptr{Cdouble}; unosc_es_nuother_8b::Ptr{Cdouble}; unosc_cc_8b::Ptr{Cdouble}
    unosc_es_nue_hep::Ptr{Cdouble}; unosc_es_nuother_hep::Ptr{Cdouble}; unosc_cc_hep::Ptr{Cdouble}
    log_e_norm::Ptr{Cdouble}
    resp_es_nue::Ptr{Cdouble}; resp_es_nuother::Ptr{Cdouble}; resp_cc::Ptr{Cdouble}
    eff_es_nue::Ptr{Cdouble}; eff_es_nuother::Ptr{Cdouble}; eff_cc::Ptr{Cdouble}
    n_bg_es::Int32; n_bg_cc::Int32
    raw_bg_es::Ptr{Cdouble}; raw_bg_cc::Ptr{Cdouble}
    bg_es_sides::Ptr{Int32}; bg_es_par::Ptr{Int32}; bg_cc_par::Ptr{Int32}
    obs_es_day::Ptr{Cdouble}; obs_es_night::Ptr{Cdouble}; obs_cc_day::Ptr{Cdouble}; obs_cc_night::Ptr{Cdouble}
    ext::Ptr{CDetectorExt}
end

struct CTables
    n_etrue::Int32; n_cosz::Int32; n_cosz_sub::Int32; n_layers::Int32; n_solar::Int32; n_det::Int32
    e_calc::Ptr{Cdouble}
    seg_start::Ptr{Int32}; seg_layer::Ptr{Int32}; seg_length::Ptr{Cdouble}
    earth_lookup::Ptr{Cdouble}; exposure_weights::Ptr{Cdouble}
    avg_ne_8b::Cdouble; avg_ne_hep::Cdouble
    solar_ne::Ptr{Cdouble}; solar_prod_8b::Ptr{Cdouble}; solar_prod_hep::Ptr{Cdouble}
    cc_xsec_norm_true::Cdouble
    det::Ptr{CDetector}
end

struct CConfig
    abi_version::Int32; device::Int32; fast::Int32; earth_uncertainty::Int32; col_major::Int32; max_batch::Int32
    night_only::Int32; reserved::Int32
end

# Parameter order fixed at create; filled by init_engine
const det_names = String[]
const n_layers = Ref(0)
const n_norms = Tuple{Int,Int}[]
const earth_unc = Ref(true)
const xsec_norm_true = Ref(1.0)
const n_cosz = Ref(0)       # coarse zenith bins / true-energy bins of the handle (shapes of the oscillation outputs)
const n_etrue = Ref(0)
const det_shapes = Dict{String,NamedTuple}()   # det_name => (nes, ncc, ncos, inclusive)

f64(x) = collect(Float64, vec(x))
i32(x) = collect(Int32, vec(x))

"""
    init_engine(E_calc, earth_paths, earth_lookup, exposure_weights, solarModel, detector_outputs,
                detector_configs; fast, earthUncertainty, cc_xsec_norm_true, device=0)

Create the handle from the globals setup.jl has just built (setup.jl:39-66 shared tables, :99-108
per-detector tables, :160-178 Asimov data and threshold indices).  Call once, after setup.jl's
detector loop; the counterpart of `nuFastOsc.init_engines` (nuFast_interface.jl:48-64).
"""
function init_engine(E_calc, earth_paths, earth_lookup, exposure_weights, solarModel, detector_outputs,
                     detector_configs; fast::Bool, earthUncertainty::Bool, cc_xsec_norm_true::Float64=1.0,
                     n_cosz_sub::Int=1, device::Int=0, max_batch::Int=0, night_only::Bool=false)
    keep = Any[]                                   # everything the C structs point into
    hold(x) = (push!(keep, x); pointer(x))

    seg_start = Int32[0]; seg_layer = Int32[]; seg_length = Float64[]
    for path in earth_paths, s in path.segments    # makePaths.jl:25-33, detector -> far surface
        push!(seg_layer, s.index); push!(seg_length, s.length)
    end
    for path in earth_paths
        push!(seg_start, seg_start[end] + length(path.segments))
    end

    empty!(det_names); empty!(n_norms)
    dets = CDetector[]
    for (name, out) in detector_outputs
        li  = out.likelihood_inputs
        u   = li.MC_no_osc; r = li.Mreco; bg = li.BG; obs = li.nObserved
        nes = length(r.eff.ES_nue); ncc = length(r.eff.CC)
        raw_es = isempty(bg.ES) ? Float64[] : reduce(vcat, f64.(bg.ES))   # component-major
        raw_cc = isempty(bg.CC) ? Float64[] : reduce(vcat, f64.(bg.CC))
        push!(det_names, name)
        push!(n_norms, (count(!=(0), bg.ES_par_counts), count(!=(0), bg.CC_par_counts)))
        let dc = detector_configs[name]
            det_shapes[name] = (nes=nes, ncc=ncc, ncos=dc.angular_reco ? r.bins.cos_scatter.bin_number : 1,
                                inclusive=dc.inclusive_analysis || dc.semi_inclusive_analysis)
        end
        # the other analysis branches of this detector (det_flags of setup.jl:140-150, tables of response.jl:111-236)
        det = detector_configs[name]
        ext = Ptr{CDetectorExt}(C_NULL)
        if det.inclusive_analysis || det.semi_inclusive_analysis || det.angular_reco
            incl = det.semi_inclusive_analysis ? 2 : (det.inclusive_analysis ? 1 : 0)
            nul = Ptr{Cdouble}(C_NULL)
            ncos = r.bins.cos_scatter.bin_number
            edges = range(r.bins.cos_scatter.min, r.bins.cos_scatter.max, length=ncos + 1)
            amask = Float64.(0.5 .* (edges[1:end-1] .+ edges[2:end]) .>= det.angular_cos_cut)   # propagation_reco.jl:98-103
            x = Ref(CDetectorExt(incl, det.angular_reco, det.angular_reco ? ncos : 0, 0,
                incl == 2 ? r.CC_split.above : 1.0, incl == 2 ? r.CC_split.below : 1.0,
                incl > 0 ? hold(f64(r.CC_inclusive)) : nul, incl > 0 ? hold(f64(r.eff.CC_incl)) : nul,
                incl == 2 ? hold(f64(r.CC_misID_nue)) : nul, incl == 2 ? hold(f64(r.CC_misID_nuother)) : nul,
                det.angular_reco ? hold(f64(r.ES.angular)) : nul, det.angular_reco ? hold(f64(r.BG.angular)) : nul,
                det.angular_reco ? hold(amask) : nul))
            push!(keep, x)
            ext = Base.unsafe_convert(Ptr{CDetectorExt}, x)
        end
        push!(dets, CDetector(
            li.ES_mode, li.CC_mode, nes, ncc, li.index_ES, li.index_CC,
            hold(f64(u.ES_nue_8B)), hold(f64(u.ES_nuother_8B)), hold(f64(u.CC_8B)),
            hold(f64(u.ES_nue_hep)), hold(f64(u.ES_nuother_hep)), hold(f64(u.CC_hep)), hold(f64(u.log_E_norm)),
            hold(f64(r.ES.nue)), hold(f64(r.ES.nuother)), hold(f64(r.CC)),          # column-major as they are
            hold(f64(r.eff.ES_nue)), hold(f64(r.eff.ES_nuother)), hold(f64(r.eff.CC)),
            length(bg.ES), length(bg.CC), hold(raw_es), hold(raw_cc),
            hold(i32(bg.sides)), hold(i32(bg.ES_par_counts)), hold(i32(bg.CC_par_counts)),
            hold(f64(obs.ES_day)), hold(f64(obs.ES_night)), hold(f64(obs.CC_day)), hold(f64(obs.CC_night)), ext))
    end
    push!(keep, dets)

    tables = CTables(
        length(E_calc) ÷ 2, length(exposure_weights), n_cosz_sub, length(earth_lookup),
        fast ? 0 : length(solarModel.n_e), length(dets),
        hold(f64(E_calc)), hold(seg_start), hold(seg_layer), hold(seg_length),
        hold(f64(earth_lookup)), hold(f64(exposure_weights)),
        solarModel.avgNeBoron, solarModel.avgNeHep,
        hold(f64(solarModel.n_e)), hold(f64(solarModel.prodFractionBoron)), hold(f64(solarModel.prodFractionHep)),
        cc_xsec_norm_true, pointer(dets))
    cfg = CConfig(ABI_VERSION, device, fast, earthUncertainty, 1, max_batch, night_only, 0)
    status = Ref{Cint}(0)
    GC.@preserve keep begin
        h = @ccall libb200.solarosc_create(Ref(tables)::Ref{CTables}, Ref(cfg)::Ref{CConfig},
                                           status::Ref{Cint})::Handle
    end
    h == C_NULL && error("solarosc_create failed (status $(status[])): " *
                         unsafe_string(@ccall libb200.solarosc_last_error(C_NULL::Handle)::Cstring))
    handle[] = h
    n_layers[] = length(earth_lookup); earth_unc[] = earthUncertainty; xsec_norm_true[] = cc_xsec_norm_true
    n_cosz[] = length(exposure_weights); n_etrue[] = length(E_calc) ÷ 2
    return h
end

shutdown_engine() = (handle[] != C_NULL && (@ccall libb200.solarosc_destroy(handle[]::Handle)::Cvoid); handle[] = C_NULL; nothing)

n_params() = Int(@ccall libb200.solarosc_n_params(handle[]::Handle)::Cint)

function check(status)
    status == 0 && return
    msg = unsafe_string(@ccall libb200.solarosc_last_error(handle[]::Handle)::Cstring)
    # status 4: a rate <= -1e-3, where poissonLogLikelihood raises (likelihood_statistical.jl:9-14)
    error("solarosc status $status: $msg")
end

# Fields of a NamedTuple whose names start with `prefix`, sorted as strings: core.jl:6-15
by_prefix(p, prefix) = [getfield(p, f) for f in sort([f for f in keys(p) if startswith(string(f), prefix)], by=string)]

"One row of the parameter matrix from the reference's NamedTuple of parameters."
function pack(p)::Vector{Float64}
    row = Float64[p.sin2_th12, p.sin2_th13, p.dm2_21, p.integrated_8B_flux, p.integrated_HEP_flux,
                  hasproperty(p, :cc_xsec_norm) ? p.cc_xsec_norm : xsec_norm_true[],   # propagation_osc.jl:185-187
                  hasproperty(p, :cc_xsec_tilt) ? p.cc_xsec_tilt : 0.0,
                  hasproperty(p, :cc_xsec_curv) ? p.cc_xsec_curv : 0.0]
    earth_unc[] && append!(row, by_prefix(p, "earth_norm_"))
    for (name, (nes, ncc)) in zip(det_names, n_norms)
        append!(row, by_prefix(p, "$(name)_ES_bg_norm_")[1:nes])
        append!(row, by_prefix(p, "$(name)_CC_bg_norm_")[1:ncc])
    end
    return row
end

pack_many(ps::AbstractVector) = reduce(hcat, pack.(ps))        # P x B, column-major == B rows of P doubles
pack_many(ps::AbstractMatrix{Float64}) = ps                      # already P x B

"""
    total_llh_batch(params) -> Vector{Float64}

lnL for every parameter point (likelihood_main.jl:30-35 applied to a batch).  `params` is a vector
of NamedTuples or a P x B matrix.  `-Inf` where an earth norm leaves [0, 2].
"""
function total_llh_batch(params; group::Int=1)
    M = pack_many(params); B = size(M, 2)
    out = Vector{Float64}(undef, B); flags = Vector{UInt32}(undef, B)
    GC.@preserve M out flags begin
        st = group == 1 ?
            (@ccall libb200.solarosc_eval_llh(handle[]::Handle, pointer(M)::Ptr{Cdouble}, Csize_t(B)::Csize_t,
                                              pointer(out)::Ptr{Cdouble}, pointer(flags)::Ptr{UInt32})::Cint) :
            (@ccall libb200.solarosc_eval_llh_grouped(handle[]::Handle, pointer(M)::Ptr{Cdouble}, Csize_t(B)::Csize_t,
                                                      Csize_t(group)::Csize_t, pointer(out)::Ptr{Cdouble},
                                                      pointer(flags)::Ptr{UInt32})::Cint)
        check(st)
    end
    return out
end

"""
    column_symbols(p) -> Vector{Union{Symbol,Nothing}}

The field of the NamedTuple `p` that feeds each column of `pack(p)` (`nothing` where `pack` falls back to a constant):
the bookkeeping the chain rule of the Dual method below needs.
"""
function column_symbols(p)
    pick(prefix) = sort([f for f in keys(p) if startswith(string(f), prefix)], by=string)
    cols = Union{Symbol,Nothing}[:sin2_th12, :sin2_th13, :dm2_21, :integrated_8B_flux, :integrated_HEP_flux,
                                 hasproperty(p, :cc_xsec_norm) ? :cc_xsec_norm : nothing,
                                 hasproperty(p, :cc_xsec_tilt) ? :cc_xsec_tilt : nothing,
                                 hasproperty(p, :cc_xsec_curv) ? :cc_xsec_curv : nothing]
    earth_unc[] && append!(cols, pick("earth_norm_"))
    for (name, (nes, ncc)) in zip(det_names, n_norms)
        append!(cols, pick("$(name)_ES_bg_norm_")[1:nes])
        append!(cols, pick("$(name)_CC_bg_norm_")[1:ncc])
    end
    return cols
end

"""
    total_llh(parameters)

Drop-in for `total_llh(parameters)` (likelihood_main.jl:30-35), including the call BAT's HMC makes of it: with
`BATContext(ad = ADSelector(:ForwardDiff))` (mcmc.jl:102-103) the log-density is evaluated on a NamedTuple of
`ForwardDiff.Dual`s.  A `ccall` cannot take duals, so that method strips them, asks the engine for the value AND the
exact gradient of the point (through the chain rendezvous when one is running, so that all chains share one launch),
and rebuilds the dual by the chain rule: partials(lnL) = sum_k dlnL/dp_k * partials(p_k).  Nothing in mcmc.jl changes:
no new ADSelector, the same `logfuncdensity(total_llh)`.  ForwardDiff evaluates in chunks of up to 12 partials, so
with P parameters the engine is called ceil(P / 12) times per gradient; `ForwardDiff.Chunk{P}()` makes that one.
"""
function total_llh(parameters::NamedTuple)
    k = findfirst(v -> v isa ForwardDiff.Dual, values(parameters))
    k === nothing && return total_llh_batch([parameters])[1]
    return total_llh_dual(parameters, typeof(values(parameters)[k]))
end

function total_llh_dual(parameters::NamedTuple, ::Type{ForwardDiff.Dual{T,V,N}}) where {T,V,N}
    vals = map(ForwardDiff.value, parameters)                     # the same NamedTuple, plain numbers
    lnl, grad = batcher[] != C_NULL ? chain_llh_and_grad(vals) : let (l, g) = total_llh_and_grad_batch([vals]); (l[1], g[:, 1]) end
    acc = ntuple(_ -> zero(V), Val(N))
    for (col, name) in enumerate(column_symbols(parameters))
        name === nothing && continue
        x = getfield(parameters, name)
        x isa ForwardDiff.Dual || continue
        px = ForwardDiff.partials(x)
        g = grad[col]
        acc = ntuple(i -> acc[i] + g * px[i], Val(N))
    end
    return ForwardDiff.Dual{T}(V(lnl), ForwardDiff.Partials(acc))
end

"""
    B200LogDensity()

`likelihood_all_samples = B200Osc.B200LogDensity()` in place of `logfuncdensity(total_llh)` (likelihood_main.jl:39):
`logdensityof(likelihood_all_samples, params)` is what llhScan.jl:97-105, mcmc.jl (through BAT) and
derive_variables_from_chain.jl call.
"""
struct B200LogDensity end
DensityInterface.DensityKind(::B200LogDensity) = DensityInterface.IsDensity()
DensityInterface.logdensityof(::B200LogDensity, parameters) = total_llh(parameters)

"""
    logdensity_and_gradient(x::AbstractVector) -> (lnL, grad)

Value and gradient on the flat parameter vector in `pack` order -- the LogDensityProblems-style entry for samplers
that take one (AdvancedHMC used directly, or a BAT context with a custom gradient operator).
"""
function logdensity_and_gradient(x::AbstractVector{<:Real})
    M = reshape(collect(Float64, x), :, 1)
    l, g = total_llh_and_grad_batch(M)
    return l[1], g[:, 1]
end

"""
    total_llh_and_grad_batch(params) -> (lnL::Vector, grad::Matrix)   # grad is P x B

Value and gradient of `total_llh` for every point: what BAT's HMC needs from the log-density in place of the
ForwardDiff pass of mcmc.jl:102-103.  Row k of `grad` belongs to column k of `pack` (see `param_names`).
"""
function total_llh_and_grad_batch(params)
    M = pack_many(params); B = size(M, 2)
    out = Vector{Float64}(undef, B); grad = Matrix{Float64}(undef, size(M, 1), B); flags = Vector{UInt32}(undef, B)
    GC.@preserve M out grad flags begin
        check(@ccall libb200.solarosc_eval_llh_grad(handle[]::Handle, pointer(M)::Ptr{Cdouble}, Csize_t(B)::Csize_t,
                                                    pointer(out)::Ptr{Cdouble}, pointer(grad)::Ptr{Cdouble},
                                                    pointer(flags)::Ptr{UInt32})::Cint)
    end
    return out, grad
end

# ---- rendezvous for BAT's chain threads ---------------------------------------------------------------------------
# BAT evaluates logdensityof(likelihood, params) one point per chain task (mcmc.jl:96-110).  solarosc_batcher_eval
# blocks until all chains still running have deposited their point, then one launch serves them.  A chain is a Julia
# TASK, and tasks migrate between threads, so the slot is not Threads.threadid(): every task draws a slot from a
# small pool the first time it evaluates, keeps it in task-local storage, and the pool reclaims the slots of finished
# tasks (BAT starts fresh tasks for every mcmc_iterate!! call).
const batcher = Ref{Ptr{Cvoid}}(C_NULL)
const slot_owner = Union{Task,Nothing}[]
const slot_lock = ReentrantLock()

function start_batcher(n_chains::Int; want_grad::Bool=true, max_wait_ms::Int=0)
    batcher[] = @ccall libb200.solarosc_batcher_create(handle[]::Handle, Cint(n_chains)::Cint, Cint(want_grad)::Cint,
                                                       Cint(max_wait_ms)::Cint)::Ptr{Cvoid}
    batcher[] == C_NULL && error("solarosc_batcher_create failed")
    lock(slot_lock) do
        empty!(slot_owner); append!(slot_owner, fill(nothing, n_chains))
    end
    return nothing
end

"0-based rendezvous slot of the calling task."
function my_slot()::Int
    tls = task_local_storage()
    haskey(tls, :b200_slot) && return tls[:b200_slot]::Int
    me = current_task()
    slot = lock(slot_lock) do
        k = findfirst(o -> o === nothing || istaskdone(o), slot_owner)
        k === nothing && error("more concurrent chain tasks than batcher slots ($(length(slot_owner)))")
        slot_owner[k] = me
        k - 1
    end
    tls[:b200_slot] = slot
    return slot
end

"lnL and gradient of ONE point from the calling chain task; blocks until the batch it joined has run."
function chain_llh_and_grad(parameters, slot::Int=my_slot())
    row = pack(parameters); lnl = Ref{Cdouble}(0.0); grad = Vector{Float64}(undef, length(row)); fl = Ref{UInt32}(0)
    GC.@preserve row grad begin
        # @threadcall would hand the wait to a libuv worker; plain @ccall keeps the Julia thread parked in C, which
        # is what the reference's NuFast wrapper does as well (nuFast_interface.jl:94-127)
        st = @ccall libb200.solarosc_batcher_eval(batcher[]::Ptr{Cvoid}, Cint(slot)::Cint, pointer(row)::Ptr{Cdouble},
                                                  lnl::Ref{Cdouble}, pointer(grad)::Ptr{Cdouble}, fl::Ref{UInt32})::Cint
        check(st)
    end
    return lnl[], grad
end

function chain_done(slot::Int=my_slot())
    @ccall libb200.solarosc_batcher_leave(batcher[]::Ptr{Cvoid}, Cint(slot)::Cint)::Cint
    lock(slot_lock) do
        slot_owner[slot + 1] = nothing
    end
    delete!(task_local_storage(), :b200_slot)
    return nothing
end
stop_batcher() = (batcher[] != C_NULL && (@ccall libb200.solarosc_batcher_destroy(batcher[]::Ptr{Cvoid})::Cvoid); batcher[] = C_NULL; nothing)

"""
    propagateSamples(unosc, resp, params, ssm, bin_edges, raw_bg, det_flags; precomputed_osc, verbose)

Same signature and 7-tuple as propagation_main.jl:56-157, every branch (exclusive, inclusive, semi-inclusive, each
with or without the angular ES reconstruction).  The table arguments are the ones the handle was created from;
`precomputed_osc` is accepted and not needed.
"""
function propagateSamples(unosc, resp, params, ssm, bin_edges, raw_bg, det_flags; precomputed_osc=nothing, verbose=false)
    d = findfirst(==(det_flags.det_name), det_names) - 1
    sh = det_shapes[det_flags.det_name]
    r = expected_rates_batch([params], d; nz=n_cosz[], nes=sh.nes, ncc=sh.ncc, ncos=sh.ncos)
    ES_day   = sh.ncos > 1 ? r.ES_day[:, :, 1]      : r.ES_day[:, 1]
    ES_night = sh.ncos > 1 ? r.ES_night[:, :, :, 1] : r.ES_night[:, :, 1]
    # CC_incl_spectrum: the 7th element (propagation_main.jl:108-146); `nothing` in the exclusive analysis
    spectrum = nothing
    if sh.inclusive
        M = pack_many([params]); spectrum = Vector{Float64}(undef, sh.nes)
        GC.@preserve M spectrum begin
            check(@ccall libb200.solarosc_eval_cc_incl_spectrum(handle[]::Handle, Cint(d)::Cint, pointer(M)::Ptr{Cdouble},
                                                                Csize_t(1)::Csize_t, pointer(spectrum)::Ptr{Cdouble})::Cint)
        end
    end
    return ES_day, r.CC_day[:, 1], ES_night, r.CC_night[:, :, 1], r.BG_ES_tot[:, 1], r.BG_CC_tot[:, 1], spectrum
end

# Rate arrays of a batch, shaped like the reference's with the batch index last: ES_day (nes, B) or, with the
# angular reconstruction, (ncos, nes, B); ES_night (nz, nes, B) or (ncos, nes, nz, B).
function expected_rates_batch(params, det::Int=0; nz::Int, nes::Int, ncc::Int, ncos::Int=1)
    M = pack_many(params); B = size(M, 2)
    ES_day   = ncos > 1 ? Array{Float64}(undef, ncos, nes, B)     : Array{Float64}(undef, nes, B)
    ES_night = ncos > 1 ? Array{Float64}(undef, ncos, nes, nz, B) : Array{Float64}(undef, nz, nes, B)
    CC_day = Array{Float64}(undef, ncc, B); CC_night = Array{Float64}(undef, nz, ncc, B)
    BG_ES = Array{Float64}(undef, nes, B);  BG_CC = Array{Float64}(undef, ncc, B)
    GC.@preserve M ES_day ES_night CC_day CC_night BG_ES BG_CC begin
        check(@ccall libb200.solarosc_eval_rates(handle[]::Handle, Cint(det)::Cint, pointer(M)::Ptr{Cdouble},
            Csize_t(B)::Csize_t, pointer(ES_day)::Ptr{Cdouble}, pointer(ES_night)::Ptr{Cdouble},
            pointer(CC_day)::Ptr{Cdouble}, pointer(CC_night)::Ptr{Cdouble}, pointer(BG_ES)::Ptr{Cdouble},
            pointer(BG_CC)::Ptr{Cdouble})::Cint)
    end
    return (ES_day=ES_day, CC_day=CC_day, ES_night=ES_night, CC_night=CC_night, BG_ES_tot=BG_ES, BG_CC_tot=BG_CC)
end

"""
    compute_shared_osc_probs(params, solarModel)

8-field NamedTuple of propagation_osc.jl:146-155 (night arrays are (n_cosz, n_Etrue)).
"""
function compute_shared_osc_probs(params, solarModel)
    nz = n_cosz[]; nT = n_etrue[]              # shapes of the handle (propagation_main.jl:47 takes no more arguments)
    M = pack_many([params])
    d8 = Vector{Float64}(undef, nT); n8 = Matrix{Float64}(undef, nz, nT)
    dh = Vector{Float64}(undef, nT); nh = Matrix{Float64}(undef, nz, nT)
    GC.@preserve M d8 n8 dh nh begin
        check(@ccall libb200.solarosc_eval_oscprobs(handle[]::Handle, pointer(M)::Ptr{Cdouble}, Csize_t(1)::Csize_t,
            pointer(d8)::Ptr{Cdouble}, pointer(n8)::Ptr{Cdouble}, pointer(dh)::Ptr{Cdouble}, pointer(nh)::Ptr{Cdouble})::Cint)
    end
    return (nue_8B_day=d8, nue_8B_night=n8, nue_hep_day=dh, nue_hep_night=nh,
            nuother_8B_day=1 .- d8, nuother_8B_night=1 .- n8, nuother_hep_day=1 .- dh, nuother_hep_night=1 .- nh)
end

"""
    total_llh_batch_multi(handles, params; group=1)

The same batch dealt over several handles -- one per GPU of the box, each created by `init_engine(...; device=k)` and
collected in `handles::Vector{Ptr{Cvoid}}` -- from this one Julia process: every GPU evaluates its own contiguous
slice and the slices' lnL land in one vector (`solarosc_eval_llh_multi`; no NCCL, no extra processes).
"""
function total_llh_batch_multi(handles::Vector{Ptr{Cvoid}}, params; group::Int=1)
    M = pack_many(params); B = size(M, 2)
    out = Vector{Float64}(undef, B); flags = Vector{UInt32}(undef, B)
    GC.@preserve handles M out flags begin
        check(@ccall libb200.solarosc_eval_llh_multi(pointer(handles)::Ptr{Ptr{Cvoid}}, Csize_t(length(handles))::Csize_t,
            pointer(M)::Ptr{Cdouble}, Csize_t(B)::Csize_t, Csize_t(group)::Csize_t, pointer(out)::Ptr{Cdouble},
            pointer(flags)::Ptr{UInt32})::Cint)
    end
    return out
end

"""
    derived_block_sums(params, weights, lref; det, nz, nes, ncc)

The replay loop of derive_variables_from_chain.jl:260-444 over a block of thinned-in chain samples, reduced on the
device: weighted rate moments (`ESday_mean`, `ESnight_m2`, ...), the centred per-bin terms' sums (`S1_*`, `S2_*`,
`Ssin2_*`, `Sdm2_*`), their per-sample totals and the scalar sums.  `lref` is the NamedTuple of per-bin terms at
the Asimov truth (`ℓref_*`, derive...:190-207).  The library's arrays are row-major `[q][1 + nz][n]`; read as
Julia arrays they are `(n, 1 + nz, q)`, so `permutedims(m[:, 2:end, q])` is the reference's `(n_cosz, n_Ereco)`
night matrix and `m[:, 1, q]` its day vector.  The caller adds the sums of successive blocks.
"""
function derived_block_sums(params, weights::Vector{Float64}, lref; det::Int=0, nz::Int, nes::Int, ncc::Int)
    M = pack_many(params); B = size(M, 2)
    # reference terms as [1 + nz][n] row-major = (n, 1 + nz) in Julia
    ref_es = hcat(lref.ES_day, permutedims(lref.ES_night)); ref_cc = hcat(lref.CC_day, permutedims(lref.CC_night))
    ES_day = Array{Float64}(undef, nes, B); CC_day = Array{Float64}(undef, ncc, B); ltot = Array{Float64}(undef, 4, B)
    mom_es = Array{Float64}(undef, nes, 1 + nz, 6); mom_cc = Array{Float64}(undef, ncc, 1 + nz, 6)
    bg_es = Array{Float64}(undef, nes, 2); bg_cc = Array{Float64}(undef, ncc, 2); sc = Vector{Float64}(undef, 21)
    GC.@preserve M weights ref_es ref_cc ES_day CC_day ltot mom_es mom_cc bg_es bg_cc sc begin
        check(@ccall libb200.solarosc_eval_derived(handle[]::Handle, Cint(det)::Cint, pointer(M)::Ptr{Cdouble}, Csize_t(B)::Csize_t,
            pointer(weights)::Ptr{Cdouble}, pointer(ref_es)::Ptr{Cdouble}, pointer(ref_cc)::Ptr{Cdouble},
            pointer(ES_day)::Ptr{Cdouble}, pointer(CC_day)::Ptr{Cdouble}, pointer(ltot)::Ptr{Cdouble},
            pointer(mom_es)::Ptr{Cdouble}, pointer(mom_cc)::Ptr{Cdouble}, pointer(bg_es)::Ptr{Cdouble},
            pointer(bg_cc)::Ptr{Cdouble}, pointer(sc)::Ptr{Cdouble})::Cint)
    end
    return (ES_day=ES_day, CC_day=CC_day, ltot=ltot, mom_ES=mom_es, mom_CC=mom_cc, bg_mom_ES=bg_es, bg_mom_CC=bg_cc, scalars=sc)
end

"""
    derived_block_sums_angular(params, weights, lref; det, nz, nes, ncc, ncos)

The same with the angular ES reconstruction (derive...:239-258, 288-296, 315-318).  The library lays the ES arrays out
row-major as `[n_cos][n_E]` (day) followed by `[n_cos][n_E][n_cosz]` (night); read as Julia arrays these are
`(n_E, n_cos)` and `(n_cosz, n_E, n_cos)`, so `permutedims` brings them to the reference's `(n_cos, n_E[, n_cosz])`.
`ES_day` is `(n_E, B)`: the day spectrum summed over the cos-scatter bins (the reference's `es_total_above` is its
`[index_ES:end]`).  Returns `mom_ES_day :: (n_cos, n_E, 6)` and `mom_ES_night :: (n_cos, n_E, n_cosz, 6)`.
"""
function derived_block_sums_angular(params, weights::Vector{Float64}, lref; det::Int=0, nz::Int, nes::Int, ncc::Int, ncos::Int)
    M = pack_many(params); B = size(M, 2)
    nd = ncos * nes; nn = nd * nz
    # (n_cos, n_E) -> memory [c][j]: permutedims to (n_E, n_cos); (n_cos, n_E, n_cosz) -> memory [c][j][z]: (n_cosz, n_E, n_cos)
    ref_es = vcat(vec(permutedims(lref.ES_day)), vec(permutedims(lref.ES_night, (3, 2, 1))))
    ref_cc = hcat(lref.CC_day, permutedims(lref.CC_night))
    ES_day = Array{Float64}(undef, nes, B); CC_day = Array{Float64}(undef, ncc, B); ltot = Array{Float64}(undef, 4, B)
    mom_es = Array{Float64}(undef, nd + nn, 6); mom_cc = Array{Float64}(undef, ncc, 1 + nz, 6)
    bg_es = Array{Float64}(undef, nes, 2); bg_cc = Array{Float64}(undef, ncc, 2); sc = Vector{Float64}(undef, 21)
    GC.@preserve M weights ref_es ref_cc ES_day CC_day ltot mom_es mom_cc bg_es bg_cc sc begin
        check(@ccall libb200.solarosc_eval_derived(handle[]::Handle, Cint(det)::Cint, pointer(M)::Ptr{Cdouble}, Csize_t(B)::Csize_t,
            pointer(weights)::Ptr{Cdouble}, pointer(ref_es)::Ptr{Cdouble}, pointer(ref_cc)::Ptr{Cdouble},
            pointer(ES_day)::Ptr{Cdouble}, pointer(CC_day)::Ptr{Cdouble}, pointer(ltot)::Ptr{Cdouble},
            pointer(mom_es)::Ptr{Cdouble}, pointer(mom_cc)::Ptr{Cdouble}, pointer(bg_es)::Ptr{Cdouble},
            pointer(bg_cc)::Ptr{Cdouble}, pointer(sc)::Ptr{Cdouble})::Cint)
    end
    day = Array{Float64}(undef, ncos, nes, 6); night = Array{Float64}(undef, ncos, nes, nz, 6)
    for q in 1:6
        day[:, :, q] = permutedims(reshape(view(mom_es, 1:nd, q), nes, ncos))
        night[:, :, :, q] = permutedims(reshape(view(mom_es, nd+1:nd+nn, q), nz, nes, ncos), (3, 2, 1))
    end
    return (ES_day=ES_day, CC_day=CC_day, ltot=ltot, mom_ES_day=day, mom_ES_night=night, mom_CC=mom_cc,
            bg_mom_ES=bg_es, bg_mom_CC=bg_cc, scalars=sc)
end

# ---- one Julia process per GPU: the library's own NCCL gather ----------------------------------------------------------
"""
    comm_unique_id() -> Vector{UInt8}      (rank 0; hand the 128 bytes to every rank, e.g. MPI.bcast)
    comm_init(id, rank, world)             (collective)
    comm_destroy()                         (collective)
    total_llh_batch_sharded(params_local, n_total; group=1) -> Vector{Float64}   (lnL of ALL n_total points on every rank)

Every rank passes the rows `shard_bounds(n_total, group, rank, world)` gives it; the per-point lnL vectors are
all-gathered device-to-device with NCCL over NVLink inside the library (solarosc_eval_llh_sharded).
"""
function comm_unique_id()
    id = Vector{UInt8}(undef, 128)
    st = GC.@preserve id @ccall libb200.solarosc_comm_unique_id(pointer(id)::Ptr{Cvoid})::Cint
    st == 0 || error("solarosc_comm_unique_id failed (status $st): is libnccl.so.2 loadable?")
    return id
end
comm_init(id::Vector{UInt8}, rank::Int, world::Int) =
    GC.@preserve id check(@ccall libb200.solarosc_comm_init(handle[]::Handle, pointer(id)::Ptr{Cvoid}, Cint(rank)::Cint, Cint(world)::Cint)::Cint)
# collective, like comm_init: every rank releases the communicator together, before the handle is destroyed (a handle destroyed
# with its communicator attached aborts it without waiting for the other ranks)
comm_destroy() = check(@ccall libb200.solarosc_comm_destroy(handle[]::Handle)::Cint)
function shard_bounds(n_total::Int, group::Int, rank::Int, world::Int)
    lo = Ref{Csize_t}(0); hi = Ref{Csize_t}(0)
    st = @ccall libb200.solarosc_shard_bounds(Csize_t(n_total)::Csize_t, Csize_t(group)::Csize_t, Cint(rank)::Cint, Cint(world)::Cint,
                                              lo::Ref{Csize_t}, hi::Ref{Csize_t})::Cint
    st == 0 || error("solarosc_shard_bounds: bad arguments")
    return Int(lo[]) + 1, Int(hi[])              # 1-based inclusive range of rows
end
function total_llh_batch_sharded(params_local, n_total::Int; group::Int=1)
    M = pack_many(params_local)
    out = Vector{Float64}(undef, n_total); flags = Vector{UInt32}(undef, n_total)
    GC.@preserve M out flags begin
        check(@ccall libb200.solarosc_eval_llh_sharded(handle[]::Handle, pointer(M)::Ptr{Cdouble}, Csize_t(n_total)::Csize_t,
            Csize_t(group)::Csize_t, pointer(out)::Ptr{Cdouble}, pointer(flags)::Ptr{UInt32})::Cint)
    end
    return out
end

end # module B200Osc
