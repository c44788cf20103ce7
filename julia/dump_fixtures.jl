#=
dump_fixtures.jl -- parity-pinning kit, Julia side.

Run INSIDE a SolarOscFitter session whose setup has completed (i.e. after `include("src/setup.jl")`, e.g. by adding
`include("path/to/dump_fixtures.jl"); dump_b200_fixtures("fixtures_escc")` at the end of src/profiling.jl or from the
REPL after `julia -t auto src/readConfig.jl config.yaml`).  It evaluates the REFERENCE's own functions

    compute_shared_osc_probs (propagation_main.jl:47-54), Osc.NumOsc.Fast.osc_prob_earth (osc.jl:448-477),
    expected_rates (likelihood_core.jl:17-35) == propagateSamples (propagation_main.jl:56-157),
    per_bin_llh and total_llh (likelihood_main.jl:30-39)

at seeded parameter points (truth, prior draws, scan corners, an out-of-bounds earth norm) and writes them, together
with EVERY constant table the engine is created from (SURVEY.md Appendix A.1), to two files:

    <prefix>.bin    raw little-endian arrays, in Julia (column-major) memory order
    <prefix>.txt    manifest, one line per array:  name dtype offset_bytes ndims dim1 dim2 ...

No HDF5 / JLD2 dependency: only Julia Base.  `python tools/check_against_julia.py <prefix>` then builds the oracle and
(on a GPU box) the engine from the dumped tables and asserts both against the dumped results at the BASELINE.json
tolerances (rates 1e-9 relative, -2lnL 1e-6 absolute, P1e 5e-12); a passing run is what turns "parity unpinned" into
"parity pinned", and the .bin/.txt pair is small enough (a few MB) to commit under tests/golden/.

NOT EXECUTED IN THIS REPOSITORY (no Julia in the build image): written against the reference's globals as setup.jl
leaves them -- E_calc, earth_paths, earth_lookup, exposure_weights, solarModel, detector_outputs, detector_configs,
N_COSZ_SUB, fast, earthUncertainty, cc_xsec_norm_true, total_llh, per_bin_llh.
=#
using Random

struct FixtureWriter
    io::IOStream
    manifest::IOStream
end

function put!(w::FixtureWriter, name::AbstractString, a::AbstractArray{T}) where {T<:Union{Float64,Int32,Int64,UInt8}}
    arr = collect(a)
    println(w.manifest, join([name, string(T), string(position(w.io)), string(ndims(arr)), string.(size(arr))...], " "))
    write(w.io, arr)
end
put!(w::FixtureWriter, name::AbstractString, x::Real) = put!(w, name, Float64[x])
put!(w::FixtureWriter, name::AbstractString, a::AbstractArray{<:Real}) = put!(w, name, Float64.(a))
put!(w::FixtureWriter, name::AbstractString, a::AbstractArray{Bool}) = put!(w, name, Int32.(a))

"Parameter NamedTuple with the fields of `base`, values replaced from `changes`."
with_values(base::NamedTuple, changes::Dict{Symbol,Float64}) =
    NamedTuple{keys(base)}(Tuple(haskey(changes, k) ? changes[k] : getfield(base, k) for k in keys(base)))

"Row of the engine's parameter matrix (include/solarosc_b200.h, the order `B200Osc.pack` uses)."
function fixture_row(p, det_names_sorted, n_norms)
    by_prefix(prefix) = [getfield(p, f) for f in sort([f for f in keys(p) if startswith(string(f), prefix)], by=string)]
    row = Float64[p.sin2_th12, p.sin2_th13, p.dm2_21, p.integrated_8B_flux, p.integrated_HEP_flux,
                  hasproperty(p, :cc_xsec_norm) ? p.cc_xsec_norm : cc_xsec_norm_true,
                  hasproperty(p, :cc_xsec_tilt) ? p.cc_xsec_tilt : 0.0, hasproperty(p, :cc_xsec_curv) ? p.cc_xsec_curv : 0.0]
    earthUncertainty && append!(row, by_prefix("earth_norm_"))
    for (name, (nes, ncc)) in zip(det_names_sorted, n_norms)
        append!(row, by_prefix("$(name)_ES_bg_norm_")[1:nes])
        append!(row, by_prefix("$(name)_CC_bg_norm_")[1:ncc])
    end
    return row
end

function dump_b200_fixtures(prefix::AbstractString; n_draws::Int=24, seed::Int=20261019)
    w = FixtureWriter(open(prefix * ".bin", "w"), open(prefix * ".txt", "w"))
    names = collect(keys(detector_outputs))
    # ---- shared tables (Appendix A.1) ------------------------------------------------------------------------------
    put!(w, "e_calc", E_calc)
    seg_start = Int32[0]; seg_layer = Int32[]; seg_length = Float64[]
    for path in earth_paths
        for s in path.segments
            push!(seg_layer, s.index); push!(seg_length, s.length)
        end
        push!(seg_start, length(seg_layer))
    end
    put!(w, "seg_start", seg_start); put!(w, "seg_layer", seg_layer); put!(w, "seg_length", seg_length)
    put!(w, "earth_lookup", earth_lookup); put!(w, "exposure_weights", vec(exposure_weights))
    put!(w, "n_cosz_sub", N_COSZ_SUB)
    put!(w, "avg_ne_8b", solarModel.avgNeBoron); put!(w, "avg_ne_hep", solarModel.avgNeHep)
    put!(w, "solar_ne", solarModel.n_e); put!(w, "solar_prod_8b", solarModel.prodFractionBoron); put!(w, "solar_prod_hep", solarModel.prodFractionHep)
    put!(w, "fast", fast ? 1.0 : 0.0); put!(w, "earth_uncertainty", earthUncertainty ? 1.0 : 0.0)
    put!(w, "cc_xsec_norm_true", cc_xsec_norm_true)
    put!(w, "n_det", length(names))
    n_norms = Tuple{Int,Int}[]
    for (d, name) in enumerate(names)
        out = detector_outputs[name]; li = out.likelihood_inputs; det = detector_configs[name]
        u = li.MC_no_osc; r = li.Mreco; bg = li.BG; obs = li.nObserved
        pre = "det$(d - 1)/"
        put!(w, pre * "name", Vector{UInt8}(name))
        put!(w, pre * "es_mode", li.ES_mode ? 1.0 : 0.0); put!(w, pre * "cc_mode", li.CC_mode ? 1.0 : 0.0)
        put!(w, pre * "index_es", li.index_ES); put!(w, pre * "index_cc", li.index_CC)
        for (k, v) in (("unosc_es_nue_8b", u.ES_nue_8B), ("unosc_es_nuother_8b", u.ES_nuother_8B), ("unosc_cc_8b", u.CC_8B),
                       ("unosc_es_nue_hep", u.ES_nue_hep), ("unosc_es_nuother_hep", u.ES_nuother_hep), ("unosc_cc_hep", u.CC_hep),
                       ("log_e_norm", u.log_E_norm))
            put!(w, pre * k, v)
        end
        put!(w, pre * "resp_es_nue", r.ES.nue); put!(w, pre * "resp_es_nuother", r.ES.nuother); put!(w, pre * "resp_cc", r.CC)   # (n_Etrue, n_Ereco)
        put!(w, pre * "eff_es_nue", r.eff.ES_nue); put!(w, pre * "eff_es_nuother", r.eff.ES_nuother); put!(w, pre * "eff_cc", r.eff.CC)
        put!(w, pre * "raw_bg_es", isempty(bg.ES) ? zeros(0, 0) : reduce(hcat, bg.ES))     # (n_Ereco, n_components)
        put!(w, pre * "raw_bg_cc", isempty(bg.CC) ? zeros(0, 0) : reduce(hcat, bg.CC))
        put!(w, pre * "bg_es_sides", Int32.(bg.sides)); put!(w, pre * "bg_es_par", Int32.(bg.ES_par_counts)); put!(w, pre * "bg_cc_par", Int32.(bg.CC_par_counts))
        put!(w, pre * "obs_es_day", obs.ES_day); put!(w, pre * "obs_es_night", obs.ES_night)
        put!(w, pre * "obs_cc_day", obs.CC_day); put!(w, pre * "obs_cc_night", obs.CC_night)
        incl = det.semi_inclusive_analysis ? 2 : (det.inclusive_analysis ? 1 : 0)
        put!(w, pre * "inclusive", incl); put!(w, pre * "angular", det.angular_reco ? 1.0 : 0.0)
        if incl > 0
            put!(w, pre * "resp_cc_incl", r.CC_inclusive); put!(w, pre * "eff_cc_incl", r.eff.CC_incl)
        end
        if incl == 2
            put!(w, pre * "resp_cc_misid_nue", r.CC_misID_nue); put!(w, pre * "resp_cc_misid_nuother", r.CC_misID_nuother)
            put!(w, pre * "cc_f_above", r.CC_split.above); put!(w, pre * "cc_f_below", r.CC_split.below)
        end
        if det.angular_reco
            nb = r.bins.cos_scatter.bin_number
            edges = range(r.bins.cos_scatter.min, r.bins.cos_scatter.max, length=nb + 1)
            put!(w, pre * "n_cos", nb); put!(w, pre * "resp_es_angular", r.ES.angular); put!(w, pre * "resp_bg_angular", r.BG.angular)
            put!(w, pre * "angular_mask", Float64.(0.5 .* (edges[1:end-1] .+ edges[2:end]) .>= det.angular_cos_cut))
        end
        push!(n_norms, (count(!=(0), bg.ES_par_counts), count(!=(0), bg.CC_par_counts)))
    end

    # ---- seeded parameter points ----------------------------------------------------------------------------------------
    # union of every detector's true parameters (shared ones coincide), as total_llh expects them (mcmc.jl:25-64)
    truth = merge((detector_outputs[n].true_params for n in names)...)
    rng = MersenneTwister(seed)
    pts = NamedTuple[truth]
    wiggle(k, v) = begin
        s = string(k)
        s == "sin2_th12" ? clamp(v + 0.04 * randn(rng), 0.2, 0.4) :
        s == "sin2_th13" ? clamp(v * (1 + 0.1 * randn(rng)), 0.0, 0.25) :
        s == "dm2_21"    ? clamp(v * (1 + 0.3 * randn(rng)), 2e-5, 2e-4) :
        startswith(s, "earth_norm_") ? clamp(1 + 0.12 * randn(rng), 0.0, 2.0) :
        (s == "cc_xsec_tilt" || s == "cc_xsec_curv") ? 0.05 * randn(rng) :
        (s == "ES_asymmetry" || s == "CC_asymmetry") ? v : v * (1 + 0.02 * randn(rng))
    end
    for _ in 1:n_draws
        push!(pts, NamedTuple{keys(truth)}(Tuple(Float64(wiggle(k, getfield(truth, k))) for k in keys(truth))))
    end
    # LLH-scan corners (llhScan.jl:35-37) and the convergence study's worst case (utils/checkOscGridConvergence.jl:40-42)
    for (s12, dm) in ((0.2, 4e-5), (0.4, 1.5e-4), (0.2, 1.5e-4), (0.4, 4e-5), (0.4, 2e-4))
        push!(pts, with_values(truth, Dict(:sin2_th12 => s12, :dm2_21 => dm)))
    end
    if earthUncertainty     # out of bounds: total_llh = -Inf (likelihood_debug.jl:1-4)
        push!(pts, with_values(truth, Dict(:earth_norm_2 => 2.5)))
    end
    put!(w, "n_points", length(pts))
    put!(w, "params", reduce(hcat, [fixture_row(p, names, n_norms) for p in pts]))          # (P, n_points)
    put!(w, "total_llh", Float64[total_llh(p) for p in pts])
    ssm = solarModel
    for (i, p) in enumerate(pts)
        i > 8 && break                                      # arrays for the first points only (file size)
        pre = "point$(i - 1)/"
        osc = compute_shared_osc_probs(p, ssm)
        put!(w, pre * "nue_8B_day", osc.nue_8B_day); put!(w, pre * "nue_8B_night", osc.nue_8B_night)      # (n_cosz, n_Etrue)
        put!(w, pre * "nue_hep_day", osc.nue_hep_day); put!(w, pre * "nue_hep_night", osc.nue_hep_night)
        for (d, name) in enumerate(names)
            li = detector_outputs[name].likelihood_inputs
            r = expected_rates(li, p; precomputed_osc=osc)
            dp = pre * "det$(d - 1)/"
            put!(w, dp * "ES_day", r.ES_day); put!(w, dp * "ES_night", r.ES_night)
            put!(w, dp * "CC_day", r.CC_day); put!(w, dp * "CC_night", r.CC_night)
            put!(w, dp * "BG_ES_tot", r.BG_ES_tot); put!(w, dp * "BG_CC_tot", r.BG_CC_tot)
        end
        if i == 1
            pb = per_bin_llh(p)                               # summed over detectors (likelihood_main.jl:37)
            put!(w, pre * "perbin_ES_day", pb.ES_day); put!(w, pre * "perbin_ES_night", pb.ES_night)
            put!(w, pre * "perbin_CC_day", pb.CC_day); put!(w, pre * "perbin_CC_night", pb.CC_night)
            mp = get_mixing_parameters(p)                     # propagation_osc.jl:30-36
            lookup = earthUncertainty ? [getfield(p, Symbol("earth_norm_$k")) for k in 1:length(earth_lookup)] .* earth_lookup : earth_lookup
            put!(w, pre * "p1e", Osc.NumOsc.Fast.osc_prob_earth(E_calc, mp, lookup, earth_paths))    # p1e[ip, ie]
        end
    end
    close(w.io); close(w.manifest)
    @info "B200 fixtures written to $(prefix).bin / .txt ($(length(pts)) points)"
end
