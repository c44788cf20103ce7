/*
 * solarosc_b200.h -- C ABI of libsolarosc_b200.so
 *
 * B200-native (sm_100a) batched engine for the per-evaluation hot path of
 * ALopezMoreno/DUNE_SolarOscFitter: oscillated event-rate prediction and
 * Poisson log-likelihood for a whole batch of parameter points per call.
 *
 * The boundary follows the reference's only FFI precedent, the NuFast-Earth
 * wrapper (reference src/oscillations/nuFast_wrapper.h:1-39 declares
 * nufast_earth_create / nufast_earth_probs / nufast_earth_destroy;
 * src/oscillations/nuFast_interface.jl:45-64,94-127 binds them with @ccall):
 * an opaque handle created once from constant tables, evaluation calls that
 * take caller-owned host buffers, and a destroy call.  Unlike that wrapper every
 * call returns a status, nothing throws across the ABI, and an error string is
 * available per handle.
 *
 * Which reference function each entry point stands in for:
 *
 *   solarosc_create        setup.jl:39-66,99-108,169-178,272-285 (tables -> LikelihoodInputs) and
 *                          nuFast_wrapper.cpp:60-87 (handle creation)
 *   solarosc_set_data      setup.jl:160-166,212-217 (Asimov / observed rates)
 *   solarosc_eval_llh      total_llh, likelihood_main.jl:30-39 -> make_likelihood,
 *                          likelihood_builder.jl:216-259 (one lnL per parameter point)
 *   solarosc_eval_llh_grouped  the same for llhScan.jl:84-116, where the n_sys x 3 points of a
 *                          grid cell share one oscillation stage (precomputed_osc,
 *                          propagation_main.jl:60-66)
 *   solarosc_eval_llh_multi  the same batch dealt over the GPUs of one box from a single host process
 *   solarosc_eval_llh_grad the value + gradient BAT's HMC asks of that log-density (mcmc.jl:102-103, where the
 *                          reference differentiates total_llh with ForwardDiff)
 *   solarosc_eval_rates    expected_rates, likelihood_core.jl:17-35 == propagateSamples,
 *                          propagation_main.jl:56-157 (all branches: exclusive, inclusive, semi-inclusive,
 *                          each with or without the angular ES reconstruction)
 *   solarosc_eval_cc_incl_spectrum  the 7th element of that tuple, propagation_main.jl:108-146
 *   solarosc_eval_perbin   make_perbin_likelihood, likelihood_builder.jl:261-333
 *   solarosc_eval_asymmetry the per-sample D/N asymmetry of derive_variables_from_chain.jl:236-258
 *   solarosc_eval_derived  the replay loop of derive_variables_from_chain.jl:260-444 over a block of chain samples
 *                          (weighted rate moments, centred per-bin lnL moments and their totals)
 *   solarosc_eval_oscprobs compute_shared_osc_probs, propagation_main.jl:47-54
 *   solarosc_eval_p1e      Osc.NumOsc.Fast.osc_prob_earth, osc.jl:448-477
 *   solarosc_poisson_llh   poissonLogLikelihood, likelihood_statistical.jl:1-28
 *
 * All arithmetic is IEEE binary64.  There is no CPU fallback: every entry point
 * needs a CUDA device of compute capability 10.0 and fails with
 * SOLAROSC_ERR_CUDA otherwise.
 */
#ifndef SOLAROSC_B200_H
#define SOLAROSC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOLAROSC_ABI_VERSION 3
#define SOLAROSC_MAX_LAYERS 8   /* Earth layers (the reference's PREM has 6) */
#define SOLAROSC_MAX_BG 8       /* background components per channel        */
#define SOLAROSC_N_FIXED_PARAMS 8

/* status codes */
enum {
    SOLAROSC_OK = 0,
    SOLAROSC_ERR_ARG = 1,      /* NULL pointer, bad size, inconsistent tables */
    SOLAROSC_ERR_CUDA = 2,     /* no sm_100 device / CUDA runtime failure      */
    SOLAROSC_ERR_ALLOC = 3,
    SOLAROSC_ERR_NEGATIVE_RATE = 4 /* some point had an expected rate <= -1e-3
                                      (the reference raises, likelihood_statistical.jl:9-14);
                                      outputs are still written, lnL = NaN for those points */
};

/* per-point flag bits written to out_flags */
enum {
    SOLAROSC_FLAG_EARTH_OOB = 1u,     /* an earth_norm outside [0,2]: lnL = -Inf (likelihood_debug.jl:1-4) */
    SOLAROSC_FLAG_NEG_WARN = 2u,      /* -1e-3 < e < 0 in some bin (reference warns)                        */
    SOLAROSC_FLAG_NEG_ERROR = 4u,     /* e <= -1e-3 in some bin (reference errors): lnL = NaN              */
    SOLAROSC_FLAG_NONFINITE = 8u      /* lnL is NaN for another reason                                      */
};

/*
 * Optional per-detector tables of the inclusive, semi-inclusive and angular-ES branches
 * (reference: response.jl:111-236 builds them, propagation_reco.jl:66-199 and propagation_main.jl:98-153 use
 * them, likelihood_builder.jl:53-78 is the angular likelihood).  2-D arrays follow col_major like the others.
 *
 * With angular != 0 every ES array gains the cos-scatter axis, in the reference's index order: ES day is
 * (n_cos, n_ereco_es), ES night (n_cos, n_ereco_es, n_cosz) -- row-major [c][j] and [c][j][z] when col_major == 0,
 * [j][c] and [z][j][c] when col_major == 1 -- for obs_es_* and for the es_day / es_night outputs of
 * solarosc_eval_rates and solarosc_eval_perbin (solarosc_eval_derived: always the row-major form, see there).
 * solarosc_eval_llh_grad serves every branch (gen_grad_kernel for handles with such a detector).
 */
typedef struct solarosc_detector_ext {
    int32_t inclusive;                  /* 0 exclusive, 1 inclusive_analysis, 2 semi_inclusive_analysis  */
    int32_t angular;                    /* det_flags.angular_reco                                        */
    int32_t n_cos;                      /* cos-scatter bins (angular != 0)                               */
    int32_t reserved;
    double cc_f_above, cc_f_below;      /* responseMatrices.CC_split (semi-inclusive; response.jl:186-187) */
    const double *resp_cc_incl;         /* [n_etrue][n_ereco_es] responseMatrices.CC_inclusive (inclusive != 0) */
    const double *eff_cc_incl;          /* [n_ereco_es]          responseMatrices.eff.CC_incl             */
    const double *resp_cc_misid_nue;    /* [n_etrue][n_ereco_cc] responseMatrices.CC_misID_nue (inclusive == 2) */
    const double *resp_cc_misid_nuother;
    const double *resp_es_angular;      /* [n_cos][n_ereco_es]   responseMatrices.ES.angular (angular != 0) */
    const double *resp_bg_angular;      /* [n_cos][n_ereco_es]   responseMatrices.BG.angular; must be zero in the bins
                                           angular_mask removes, as response.jl:119-123 builds it (checked at create) */
    const double *angular_mask;         /* [n_cos] 1.0 where the bin centre >= angular_cos_cut, else 0.0
                                           (propagation_reco.jl:98-105)                                  */
} solarosc_detector_ext;

/*
 * Per-detector constant tables (reference: unoscillatedSample.jl:86-116, response.jl:111-141,
 * backgrounds.jl:53-137, setup.jl:160-178).  2-D arrays are indexed [a][b] with b fastest when
 * solarosc_config.col_major == 0 (C / numpy) and with a fastest when col_major == 1 (Julia
 * arrays passed as they lie in memory).
 */
typedef struct solarosc_detector {
    int32_t es_mode, cc_mode;           /* det.ES_mode / det.CC_mode                                  */
    int32_t n_ereco_es, n_ereco_cc;     /* reconstructed-energy bins per channel                      */
    int32_t index_es, index_cc;         /* 1-based first bin entering the likelihood (setup.jl:169-178) */
    /* unoscillated sample, each [n_etrue] */
    const double *unosc_es_nue_8b, *unosc_es_nuother_8b, *unosc_cc_8b;
    const double *unosc_es_nue_hep, *unosc_es_nuother_hep, *unosc_cc_hep;
    const double *log_e_norm;
    /* responses [n_etrue][n_ereco] (row = true-energy bin) and efficiencies [n_ereco] */
    const double *resp_es_nue, *resp_es_nuother, *resp_cc;
    const double *eff_es_nue, *eff_es_nuother, *eff_cc;
    /* raw background components (propagation_bg.jl:36-83) */
    int32_t n_bg_es, n_bg_cc;
    const double *raw_bg_es;            /* [n_bg_es][n_ereco_es] (component-major, both layouts)      */
    const double *raw_bg_cc;            /* [n_bg_cc][n_ereco_cc]                                      */
    const int32_t *bg_es_sides;         /* [n_bg_es] 0 long faces, 1 end caps, else total            */
    const int32_t *bg_es_par;           /* [n_bg_es] ES_par_counts: != 0 -> scaled by a nuisance      */
    const int32_t *bg_cc_par;           /* [n_bg_cc] CC_par_counts                                    */
    /* observed / Asimov data; all four may be NULL at create and supplied by solarosc_set_data */
    const double *obs_es_day;           /* [n_ereco_es]           */
    const double *obs_es_night;         /* [n_cosz][n_ereco_es]   */
    const double *obs_cc_day;           /* [n_ereco_cc]           */
    const double *obs_cc_night;         /* [n_cosz][n_ereco_cc]   */
    const solarosc_detector_ext *ext;   /* NULL: exclusive, non-angular analysis                      */
} solarosc_detector;

/* Shared constant tables */
typedef struct solarosc_tables {
    int32_t n_etrue;                    /* nBins_Etrue; the oscillation grid has 2*n_etrue energies   */
    int32_t n_cosz;                     /* coarse zenith bins                                         */
    int32_t n_cosz_sub;                 /* N_COSZ_SUB (earthProfile.jl:165); n_paths = n_cosz*n_cosz_sub */
    int32_t n_layers;                   /* Earth layers, <= SOLAROSC_MAX_LAYERS                       */
    int32_t n_solar;                    /* solar-model radii (0 allowed when fast != 0)               */
    int32_t n_det;                      /* detectors summed by total_llh                              */
    const double *e_calc;               /* [2*n_etrue] GeV (setup.jl:39-41, propagation_osc.jl:8)     */
    const int32_t *seg_start;           /* [n_paths+1] CSR offsets into seg_layer/seg_length          */
    const int32_t *seg_layer;           /* [n_seg] 1-based layer id, detector -> far surface          */
    const double *seg_length;           /* [n_seg] km                                                 */
    const double *earth_lookup;         /* [n_layers] mean potential per layer (setup.jl:53-66)       */
    const double *exposure_weights;     /* [n_cosz] (exposure.jl:40-43)                               */
    double avg_ne_8b, avg_ne_hep;       /* solarModel.jl:19-20                                        */
    const double *solar_ne, *solar_prod_8b, *solar_prod_hep; /* [n_solar] (fast == 0 only)            */
    double cc_xsec_norm_true;           /* unused placeholder for hosts that fix the CC cross-section */
    const solarosc_detector *det;       /* [n_det]                                                    */
} solarosc_tables;

typedef struct solarosc_config {
    int32_t abi_version;                /* must be SOLAROSC_ABI_VERSION                               */
    int32_t device;                     /* CUDA device ordinal                                        */
    int32_t fast;                       /* fastFit (config.yaml:20): 1 = mean density, 0 = production-region integral */
    int32_t earth_uncertainty;          /* earth_potential_uncertainties: 1 = earth_norm_k columns present */
    int32_t col_major;                  /* 0 = C/numpy layout, 1 = Julia layout for 2-D inputs/outputs */
    int32_t max_batch;                  /* points staged per internal chunk (0 = default 4096)        */
    int32_t night_only;                 /* 1 = likelihoods (lnL, gradient) sum the night rows only: the day vector of
                                           every channel stays out of the Poisson sum (BASELINE config 4, the night-only
                                           Earth-density study); rates and per-bin outputs are unaffected         */
    int32_t reserved;                   /* must be 0                                                  */
} solarosc_config;

typedef struct solarosc_handle solarosc_handle;

/*
 * Parameter matrix: params[b*P + k], P = solarosc_n_params(h), columns
 *   0 sin2_th12  1 sin2_th13  2 dm2_21  3 integrated_8B_flux  4 integrated_HEP_flux
 *   5 cc_xsec_norm  6 cc_xsec_tilt  7 cc_xsec_curv
 *   then earth_norm_1..n_layers (only if earth_uncertainty),
 *   then for each detector d: {det}_ES_bg_norm_1..nES_d, {det}_CC_bg_norm_1..nCC_d where
 *   nES_d / nCC_d count the non-zero entries of bg_es_par / bg_cc_par (the families the reference
 *   gathers with _params_by_prefix, core.jl:6-15).  Both channels' columns are always present,
 *   whatever es_mode / cc_mode say.
 */
solarosc_handle *solarosc_create(const solarosc_tables *tables, const solarosc_config *config, int *status);
void solarosc_destroy(solarosc_handle *h);
const char *solarosc_last_error(const solarosc_handle *h); /* h == NULL: error of the last failed create */
int solarosc_n_params(const solarosc_handle *h);
int solarosc_abi_version(void);

int solarosc_set_data(solarosc_handle *h, int det, const double *obs_es_day, const double *obs_es_night,
                      const double *obs_cc_day, const double *obs_cc_night);

/* lnL per point (multiply by -2 for the -2lnL of BASELINE.json); out_flags may be NULL. */
int solarosc_eval_llh(solarosc_handle *h, const double *params, size_t n_points, double *out_lnl,
                      uint32_t *out_flags);

/* As above; points [g*group, (g+1)*group) use the oscillation columns (0..2 and earth_norm_*) of
 * point g*group.  n_points must be a multiple of group. */
int solarosc_eval_llh_grouped(solarosc_handle *h, const double *params, size_t n_points, size_t group,
                              double *out_lnl, uint32_t *out_flags);

/* Several GPUs from one process: hs[k] are handles created from the same tables on different devices
 * (solarosc_config.device).  The n_points rows are dealt to the handles in contiguous slices of whole groups, every
 * slice is evaluated concurrently by its own host thread, and each slice's lnL is copied straight into out_lnl --
 * the gather of the per-point vector that a multi-process run does with NCCL (bench.py, sharding.py) happens here
 * through the host array the caller wants the results in anyway.  group = 1 for independent points. */
int solarosc_eval_llh_multi(solarosc_handle *const *hs, size_t n_handles, const double *params, size_t n_points,
                            size_t group, double *out_lnl, uint32_t *out_flags);

/*
 * One process per GPU: every rank creates its handle on its own device, the ranks share an
 * NCCL communicator, each evaluates its contiguous slice of the batch and the per-point lnL / flag vectors are
 * all-gathered device-to-device over NVLink -- the only collective of the path, 12 bytes per point.  NCCL is bound at run
 * time (dlopen of libnccl.so.2); without it these calls return SOLAROSC_ERR_CUDA and everything else works.
 *
 *   solarosc_comm_unique_id   rank 0: 128 bytes (an ncclUniqueId) to hand to every rank by the host's own means
 *                             (MPI.jl bcast, a shared file, torch.distributed ...)
 *   solarosc_comm_init        collective: every rank, same id, its rank and the world size
 *   solarosc_shard_bounds     rows [lo, hi) of a rank: whole groups, dealt as evenly as possible, earlier ranks first
 *   solarosc_eval_llh_sharded every rank passes ITS rows (params_local = the rows solarosc_shard_bounds gives it, host
 *                             buffer) and receives the lnL / flags of ALL n_total points
 *   solarosc_eval_llh_sharded_device  the same with device pointers on `stream`, returning without synchronising
 */
#define SOLAROSC_COMM_ID_BYTES 128
int solarosc_comm_unique_id(void *id_out);
int solarosc_comm_init(solarosc_handle *h, const void *id, int rank, int world_size);
int solarosc_comm_destroy(solarosc_handle *h);
int solarosc_shard_bounds(size_t n_total, size_t group, int rank, int world_size, size_t *lo, size_t *hi);
int solarosc_eval_llh_sharded(solarosc_handle *h, const double *params_local, size_t n_total, size_t group,
                              double *out_lnl, uint32_t *out_flags);
int solarosc_eval_llh_sharded_device(solarosc_handle *h, const double *d_params_local, size_t n_total, size_t group,
                                     double *d_lnl_all, uint32_t *d_flags_all, void *stream);

/* lnL and d lnL / d params for every point: out_grad[b*P + k] = d lnL_b / d params[b][k], every column exact to
 * rounding like the reference's ForwardDiff duals (mcmc.jl:102-103).  Fluxes, cross-section and background norms are
 * differentiated in reverse through the Poisson sum and the response fold (grad_kernel); the oscillation columns
 * (0..2 and earth_norm_*) by an adjoint sweep through the layered Earth propagation and the analytic derivative of the
 * MSW factors (adjoint_kernel; against 40-digit differentiation of the reference formula: 1e-13 relative,
 * tests/test_adjoint_highprec.py).  sin2_th12 and sin2_th13 are clamped to [0, 1] by the reference
 * (propagation_osc.jl:30-36): their columns are 0 outside the open interval.  Columns the likelihood does not depend
 * on ({det}_ES_bg_norm_2.., which the reference never reads) get 0, as does every column of a point whose earth
 * norms are out of bounds (lnL = -Inf).  Every analysis branch (exclusive, inclusive, semi-inclusive, each with or
 * without the angular ES reconstruction; a handle may mix them); night_only handles are refused. */
int solarosc_eval_llh_grad(solarosc_handle *h, const double *params, size_t n_points, double *out_lnl,
                           double *out_grad, uint32_t *out_flags);

/* Expected rates of detector det.  Outputs (any may be NULL), batch index slowest:
 *   es_day [n][n_ereco_es], es_night [n][n_cosz][n_ereco_es] (col_major: [n][n_ereco_es][n_cosz]),
 *   cc_day, cc_night likewise, bg_es [n][n_ereco_es], bg_cc [n][n_ereco_cc]. */
int solarosc_eval_rates(solarosc_handle *h, int det, const double *params, size_t n_points, double *es_day,
                        double *es_night, double *cc_day, double *cc_night, double *bg_es, double *bg_cc);

/* CC_incl_spectrum, the 7th element of propagateSamples' tuple (propagation_main.jl:108-146; setup.jl:235-243 uses it
 * to print the ES / CC / background split of the inclusive data): the inclusive-CC part of the ES arrays summed over the
 * day vector and every night row (and the cos-scatter bins), times CC_split.above in the semi-inclusive analysis.
 * out [n][n_ereco_es].  Detectors of the exclusive analysis have none (the reference returns `nothing`):
 * SOLAROSC_ERR_ARG. */
int solarosc_eval_cc_incl_spectrum(solarosc_handle *h, int det, const double *params, size_t n_points, double *out);

/* Per-bin Poisson terms of detector det, shaped like the data (zeros outside idx..n-1, -Inf
 * everywhere for a point whose earth norms are out of bounds). */
int solarosc_eval_perbin(solarosc_handle *h, int det, const double *params, size_t n_points, double *es_day,
                         double *es_night, double *cc_day, double *cc_night);

/* Day-night asymmetry 2(D - N)/(D + N) of detector det per point and channel, from the background-subtracted
 * rate totals above threshold (derive_variables_from_chain.jl:236-258), reduced on the device: 16 bytes per
 * point come back instead of the 49 kB of rate arrays.  Either output may be NULL. */
int solarosc_eval_asymmetry(solarosc_handle *h, int det, const double *params, size_t n_points, double *asym_es,
                            double *asym_cc);

/* Derived-mode reductions over a block of (thinned-in) chain samples of detector det, accumulated on the device
 * (derive_variables_from_chain.jl:260-444): O(bins) numbers come back instead of 49 kB of rate arrays per sample.
 * The caller adds the block sums of successive calls.  Array outputs are row-major whatever col_major says; any
 * output may be NULL.
 *   weights   [n]  chain weights w
 *   lref_es / lref_cc  [1 + n_cosz][n_ereco] reference per-bin terms (row 0 = day) the per-bin terms are centred on
 *             (the reference uses the per-bin terms at the Asimov truth, derive...:190-207); NULL = zeros
 *   es_day / cc_day    [n][n_ereco] day rates per sample (the posterior-predictive sample lists, derive...:315-332)
 *   ltot      [n][4]  totals of the centred per-bin terms: ES night, ES day, CC night, CC day
 *   mom_es / mom_cc    [6][1 + n_cosz][n_ereco]: sum w r, sum w r^2 (bins from the threshold on, zero below),
 *             sum w l, sum w l^2, sum w l sin2_th12, sum w l dm2_21 with l the centred per-bin term
 *   bg_mom_es / bg_mom_cc  [2][n_ereco]: sum w (BG/2), sum w (BG/2)^2
 *   scalars   [21]: sum w, sum w th1, sum w th1^2, sum w th2, sum w th2^2 (th1 = sin2_th12, th2 = dm2_21), then for
 *             each group of ltot: sum w L, sum w L^2, sum w L th1, sum w L th2
 * With the angular ES reconstruction (derive...:239-258, 288-296, 315-318) the ES arrays carry the cos-scatter axis:
 * lref_es and each of the six mom_es blocks hold n_cos n_ereco_es (1 + n_cosz) doubles, the day part [n_cos][n_ereco_es]
 * followed by the night part [n_cos][n_ereco_es][n_cosz] (the reference's index order, row-major); es_day stays
 * [n][n_ereco_es]: the day rates summed over the cos-scatter bins, which is what the reference lists per sample. */
int solarosc_eval_derived(solarosc_handle *h, int det, const double *params, size_t n_points, const double *weights,
                          const double *lref_es, const double *lref_cc, double *es_day, double *cc_day, double *ltot,
                          double *mom_es, double *mom_cc, double *bg_mom_es, double *bg_mom_cc, double *scalars);

/* nu_e survival probabilities at analysis resolution: day [n][n_etrue], night [n][n_cosz][n_etrue]
 * (col_major: [n][n_etrue][n_cosz]); nu_other = 1 - nu_e. */
int solarosc_eval_oscprobs(solarosc_handle *h, const double *params, size_t n_points, double *day_8b,
                           double *night_8b, double *day_hep, double *night_hep);

/* Earth-matter P(nu_1 -> nu_e) at oscillation resolution: [n][n_paths][2*n_etrue]
 * (col_major: [n][2*n_etrue][n_paths], i.e. the reference's p1e[ip, ie]). */
int solarosc_eval_p1e(solarosc_handle *h, const double *params, size_t n_points, double *p1e);

/* Device-resident variants: d_params / d_lnl / d_flags are device pointers on the handle's device,
 * work is enqueued on `stream` (a cudaStream_t passed as void*, NULL = the handle's own stream) and
 * the call returns without synchronising. */
int solarosc_eval_llh_device(solarosc_handle *h, const double *d_params, size_t n_points, size_t group,
                             double *d_lnl, uint32_t *d_flags, void *stream);

/* Stand-alone Poisson log-likelihood of n_vec independent (expected, measured) vector pairs of
 * length len each, host buffers: out[v] = poissonLogLikelihood(e[v,:], m[v,:]). */
int solarosc_poisson_llh(solarosc_handle *h, const double *expected, const double *measured, size_t n_vec,
                         size_t len, double *out, uint32_t *out_flags);

/*
 * Rendezvous batching for hosts that evaluate one point per thread (BAT's chains on Julia threads call
 * logdensityof(likelihood, params) one point at a time, mcmc.jl:96-110, likelihood_main.jl:37-39): each of
 * n_participants threads owns a slot and calls solarosc_batcher_eval with its point; the call blocks until every
 * participant that has not left is waiting (or until max_wait_ms > 0 has elapsed for some waiter), then ONE
 * batched evaluation serves them all.  want_grad selects solarosc_eval_llh_grad (out_grad [P]) over
 * solarosc_eval_llh.  The return value is the status of the batched call.  solarosc_batcher_leave retires a slot
 * (a chain that has finished) so the others stop waiting for it.  solarosc_batcher_create_with takes any
 * evaluator with the same contract (tests drive the rendezvous logic with a CPU stand-in).
 */
typedef struct solarosc_batcher solarosc_batcher;
typedef int (*solarosc_batch_fn)(void *ctx, const double *params, size_t n_points, double *out_lnl, double *out_grad,
                                 uint32_t *out_flags);
solarosc_batcher *solarosc_batcher_create(solarosc_handle *h, int n_participants, int want_grad, int max_wait_ms);
solarosc_batcher *solarosc_batcher_create_with(solarosc_batch_fn fn, void *ctx, int n_params, int n_participants,
                                               int want_grad, int max_wait_ms);
int solarosc_batcher_eval(solarosc_batcher *b, int slot, const double *params, double *out_lnl, double *out_grad,
                          uint32_t *out_flag);
int solarosc_batcher_leave(solarosc_batcher *b, int slot);
int solarosc_batcher_stats(solarosc_batcher *b, unsigned long long *n_batches, unsigned long long *n_points);
void solarosc_batcher_destroy(solarosc_batcher *b);

/* Per-kernel device timing.  When enabled, eval_llh / eval_llh_grouped / eval_llh_device record CUDA
 * events around each kernel on the launching stream; solarosc_last_timing waits for the last one and
 * returns the summed durations [ms] of the oscillation kernel and of the rate+likelihood kernel of
 * the last such call.  Its return value is the number of kernels that call launched (always
 * counted, timing or not). */
int solarosc_set_timing(solarosc_handle *h, int enabled);
int solarosc_last_timing(solarosc_handle *h, float *ms_osc, float *ms_rates);
/* 1 if the last likelihood call ran the fused kernel (llh_kernel: ms_osc is its time, ms_rates ~ 0),
 * 0 if it ran osc_kernel + rates_kernel. */
int solarosc_last_fused(const solarosc_handle *h);

/* Measured FP64 FMA throughput of the handle's device in TFLOP/s (dependent-chain DFMA probe): the
 * denominator of the FP64 roofline, which MEASURED_PEAKS.json does not carry. */
int solarosc_measure_fp64_peak(solarosc_handle *h, double *tflops);

/*
 * Set-up stage: the constant tables solarosc_create consumes, built from the MC samples and the Earth model the way
 * the reference's Julia set-up does (response.jl:3-141, backgrounds.jl:37-140, histHelpers.jl:11-78,
 * makePaths.jl:48-294, setup.jl:52-66).  Host functions (one-time work); integer bin indices are bit-exact
 * (left-closed bins, searchsortedlast), weights are summed in file order.  edges == NULL means n_bins equal bins on
 * [lo, hi]; otherwise n_bins + 1 explicit edges (the reference pushes the outer reco edges to -Inf / +Inf while
 * still testing the range [lo, hi], response.jl:47 vs :98-104).
 *
 *   solarosc_setup_bin_index         out[i] = searchsortedlast(edges, v[i]) (0 below the first edge)
 *   solarosc_setup_bin_index_device  the same on the GPU of handle h (one thread per event), host buffers
 *   solarosc_setup_histogram         create_histogram: keep (NULL = all) selects events, w (NULL = 1) weighs them
 *   solarosc_setup_response_matrix   create_response_matrix: out [nbx][nby] row-normalised; xi / yi (optional) receive
 *                                    every event's 1-based (true, reco) bin, 0 for events outside the range
 *   solarosc_setup_earth_paths       make_potential_for_integrand + get_avg_densities for the given zenith cosines:
 *                                    CSR segment tables (layer ids 1-based, detector -> far surface) and the mean
 *                                    potential per layer (lookup, at least n_layers_expected entries)
 */
int solarosc_setup_bin_index(const double *edges, size_t n_edges, const double *v, size_t n, int32_t *out);
int solarosc_setup_bin_index_device(solarosc_handle *h, const double *edges, size_t n_edges, const double *v, size_t n,
                                    int32_t *out);
int solarosc_setup_histogram(const double *v, const double *w, const uint8_t *keep, size_t n, int32_t n_bins, double lo,
                             double hi, const double *edges, int32_t normalise, double *out);
int solarosc_setup_response_matrix(const double *x, const double *y, size_t n, int32_t nbx, double xlo, double xhi,
                                   const double *xedges, int32_t nby, double ylo, double yhi, const double *yedges,
                                   double *out, int32_t *xi, int32_t *yi);
int solarosc_setup_earth_paths(const double *radius, const double *potential, size_t n_model, const double *cosz,
                               size_t n_paths, int32_t n_layers_expected, int32_t *seg_start, int32_t *seg_layer,
                               double *seg_length, size_t seg_capacity, double *lookup, int32_t *n_layers_found);

#ifdef __cplusplus
}
#endif
#endif /* SOLAROSC_B200_H */
