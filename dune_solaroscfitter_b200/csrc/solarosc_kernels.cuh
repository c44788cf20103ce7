// solarosc_kernels.cuh -- the two batched kernels of the hot path.
//
//   osc_kernel    K1 + K2: adiabatic MSW day probability (mean density or production-region
//                 integral) and Earth-matter P1e through the layered path table, fused with the
//                 IYSW night formula and the (N_COSZ_SUB x 2) block average.  One CTA per
//                 (parameter point, chunk of coarse energy bins).
//   rates_kernel  K3 + K4 + K5: flux x cross-section x exposure weighting, E_true -> E_reco fold
//                 for ES nu_e / ES nu_other / CC, background addition and the soft-floored Poisson
//                 log-likelihood reduced with warp shuffles.  One CTA per parameter point.
//
// Reference code each stage stands in for is cited at the stage.
#pragma once
#include "solarosc_device.cuh"

namespace solarosc {

constexpr int kOscThreads = 256;
constexpr int kRatesThreads = 256;

struct OscLaunch {
    int n_points;     // oscillation points in this launch
    int n_chunks;     // energy chunks per point
    int cb;           // coarse energy bins per chunk
    int param_stride; // doubles between consecutive points' parameter rows
    int point_step;   // parameter row of oscillation point g is g*point_step (grouped evaluation)
};

__host__ __device__ inline size_t osc_smem_bytes(int n_layers, int cb) {
    return sizeof(double) * ((size_t)11 * n_layers * 2 * cb + (size_t)4 * 2 * cb);
}

// ------------------------------------------------------------------------------------------
// K1 + K2
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kOscThreads)
osc_kernel(const DevTables T, const OscLaunch L, const double* __restrict__ params,
           double* __restrict__ osc_out /* [n_points][2][1+n_cosz][n_etrue] */,
           double* __restrict__ p1e_out /* [n_points][n_paths][n_e] or nullptr */) {
    extern __shared__ double smem[];
    const int g = blockIdx.x / L.n_chunks;
    const int chunk = blockIdx.x - g * L.n_chunks;
    const int i0 = chunk * L.cb;                         // first coarse energy bin of this CTA
    const int cb = min(L.cb, T.n_etrue - i0);            // coarse bins handled here
    const int ne = 2 * cb;                               // fine energies handled here
    const int nl = T.n_layers;
    const int cbs = L.cb;                                // table stride (full chunk size)
    double* tbl = smem;                                  // [11][nl][2][cbs]
    double* msw = smem + (size_t)11 * nl * 2 * cbs;      // [2 proc][2 what][2][cbs]
    const double* p = params + (size_t)g * L.point_step * L.param_stride;

    // mixing parameters (reference propagation_osc.jl:30-36: theta = asin(sqrt(clamp(sin2, 0, 1))))
    const double s12sq = fmin(fmax(p[0], 0.0), 1.0), s13sq = fmin(fmax(p[1], 0.0), 1.0);
    const double dm21 = p[2];
    const double c13sq = 1.0 - s13sq, c12sq = 1.0 - s12sq;
    const double c13 = sqrt(c13sq);
    const double u1 = sqrt(c12sq) * c13, u2 = sqrt(s12sq) * c13, u3 = sqrt(s13sq);
    const double p0 = c12sq * c13sq;                     // P_0 = c12^2 c13^2 (osc.jl:196)
    const double c2 = 1.0 - 2.0 * s12sq;                 // cos(2 theta12)

    // ---- phase 1a: one eigen-system per (fine energy, layer)  (osc.jl:466-470, 370-388) --------
    for (int t = threadIdx.x; t < ne * nl; t += blockDim.x) {
        const int k = t / ne, el = t - k * ne;
        const double e = T.e_calc[2 * i0 + el];
        const double norm = T.earth_unc ? p[SOLAROSC_N_FIXED_PARAMS + k] : 1.0;
        const double a = norm * T.earth_lookup[k] * e;   // lookup = earth_norm .* earth_lookup (propagation_osc.jl:60)
        double q[9], w2, w3;
        eig_rank_one(dm21, kDm31, a, u1, u2, u3, q, w2, w3);
        const int slot = (el & 1) * cbs + (el >> 1);
        const double f = kPhase / e;                      // phase = c * (l/e) * lambda (osc.jl:361)
#pragma unroll
        for (int c = 0; c < 9; ++c) tbl[((size_t)c * nl + k) * 2 * cbs + slot] = q[c];
        tbl[((size_t)9 * nl + k) * 2 * cbs + slot] = f * w2;
        tbl[((size_t)10 * nl + k) * 2 * cbs + slot] = f * w3;
    }
    // ---- phase 1b: MSW day probability and cos(2 theta_m) per (process, fine energy) ------------
    if (T.fast) {  // osc.jl:185-207 with n_e = production-weighted mean density
        for (int t = threadIdx.x; t < 2 * ne; t += blockDim.x) {
            const int proc = t / ne, el = t - proc * ne;
            const double e = T.e_calc[2 * i0 + el];
            const double c2m = msw_c2m(e, T.avg_ne[proc], dm21, c13sq, c2);
            const int slot = (el & 1) * cbs + (el >> 1);
            msw[(size_t)(proc * 2 + 0) * 2 * cbs + slot] = fma(0.5 * c13sq * c13sq, fma(c2, c2m, 1.0), s13sq * s13sq);
            msw[(size_t)(proc * 2 + 1) * 2 * cbs + slot] = c13sq * c2m;
        }
    } else {  // osc.jl:139-182: the radius loop is linear in (P_day, cos2theta_m), so both collapse to
              // production-weighted means; one warp per (process, energy), lanes over radii
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
        for (int t = warp; t < 2 * ne; t += nw) {
            const int proc = t / ne, el = t - proc * ne;
            const double e = T.e_calc[2 * i0 + el];
            const double* w = T.solar_w[proc];
            double acc = 0.0;
            for (int r = lane; r < T.n_solar; r += 32) acc = fma(w[r], msw_c2m(e, T.solar_ne[r], dm21, c13sq, c2), acc);
            acc = warp_sum(acc);
            if (lane == 0) {
                const double c2m = acc / T.solar_wsum[proc];
                const int slot = (el & 1) * cbs + (el >> 1);
                msw[(size_t)(proc * 2 + 0) * 2 * cbs + slot] = fma(0.5 * c13sq * c13sq, fma(c2, c2m, 1.0), s13sq * s13sq);
                msw[(size_t)(proc * 2 + 1) * 2 * cbs + slot] = c13sq * c2m;
            }
        }
    }
    __syncthreads();

    const int nrow = 1 + T.n_cosz;
    double* out8 = osc_out + (size_t)g * 2 * nrow * T.n_etrue;
    double* outh = out8 + (size_t)nrow * T.n_etrue;

    // ---- day rows: block_average(prob_day, 2)  (propagation_osc.jl:133-134) -----------------------
    for (int j = threadIdx.x; j < cb; j += blockDim.x) {
        out8[i0 + j] = 0.5 * (msw[j] + msw[cbs + j]);
        outh[i0 + j] = 0.5 * (msw[(size_t)2 * 2 * cbs + j] + msw[(size_t)2 * 2 * cbs + cbs + j]);
    }

    // ---- phase 2: Earth propagation, one item = (coarse zenith row, coarse energy bin) -----------
    // Lanes of a warp walk the same path at neighbouring energies, so the layer loop is
    // warp-uniform; rows are visited longest-first so the tail of the item loop is the cheap rows.
    const double inv_block = 1.0 / (2.0 * T.n_sub);
    const int n_items = T.n_cosz * cb;
    for (int item = threadIdx.x; item < n_items; item += blockDim.x) {
        const int zr = item / cb, j = item - zr * cb;
        const int z = T.row_order[zr];
        double acc8 = 0.0, acch = 0.0;
        for (int sub = 0; sub < T.n_sub; ++sub) {
            const int path = z * T.n_sub + sub;
            const int s0 = T.path_start[path], s1 = T.path_start[path + 1];
#pragma unroll
            for (int par = 0; par < 2; ++par) {
                const int slot = par * cbs + j;
                // state in the vacuum mass basis, starts as |nu_e> = u
                double ar = u1, br = u2, cr = u3, ai = 0.0, bi = 0.0, ci = 0.0;
                for (int s = s0; s < s1; ++s) {          // osc_reduce_from_lookup (osc.jl:436-446)
                    const int k = T.app_layer[s];
                    const double len = T.app_len[s];
                    earth_step(tbl + (size_t)k * 2 * cbs + slot, (size_t)nl * 2 * cbs, len, ar, ai, br, bi, cr, ci);
                }
                const double p1e = fma(ar, ar, ai * ai);  // |<nu_1|X|nu_e>|^2 (osc.jl:444)
                if (p1e_out) p1e_out[((size_t)g * T.n_paths + path) * T.n_e + 2 * (i0 + j) + par] = p1e;
                // IYSW night probability (osc.jl:204): P_day + c13^2 cos2theta_m (P1e - P0)
                const double d = p1e - p0;
                acc8 += fma(msw[(size_t)1 * 2 * cbs + slot], d, msw[slot]);
                acch += fma(msw[(size_t)3 * 2 * cbs + slot], d, msw[(size_t)2 * 2 * cbs + slot]);
            }
        }
        out8[(size_t)(1 + z) * T.n_etrue + i0 + j] = acc8 * inv_block;   // block_average (propagation_osc.jl:136-137)
        outh[(size_t)(1 + z) * T.n_etrue + i0 + j] = acch * inv_block;
    }
}

// ------------------------------------------------------------------------------------------
// K3 + K4 + K5
// ------------------------------------------------------------------------------------------
enum RatesMode { kModeLlh = 0, kModeRates = 1, kModePerbin = 2 };

struct RatesLaunch {
    int n_points;
    int group;         // points per shared oscillation stage
    int param_stride;
    int mode;          // RatesMode
    int det_only;      // -1: all detectors (likelihood), else the one detector whose arrays are written
    long long s_row_es, s_j_es, s_row_cc, s_j_cc;  // output strides of the night arrays (layout)
};

struct RatesOut {
    double* lnl; uint32_t* flags;
    double *es_day, *es_night, *cc_day, *cc_night, *bg_es, *bg_cc;
};

__host__ __device__ inline size_t rates_smem_bytes(int n_etrue, int n_cosz, int max_ereco) {
    return sizeof(double) * ((size_t)3 * (1 + n_cosz) * n_etrue + (size_t)2 * max_ereco + 64);
}

// soft-floored Poisson term (reference likelihood_statistical.jl:16-24)
__device__ __forceinline__ double poisson_term(double raw_e, double m, double logm, uint32_t& fl) {
    if (raw_e < 0.0) fl |= (raw_e > -1e-3) ? SOLAROSC_FLAG_NEG_WARN : SOLAROSC_FLAG_NEG_ERROR;
    const double d = raw_e - 1e-9;
    const double e = 0.5 * (raw_e + 1e-9 + sqrt(fma(d, d, 1e-18)));
    return (m > 0.0) ? (e - m + m * (logm - log(e))) : e;
}

// per-bin diagnostic term (reference likelihood_statistical.jl:31-57)
__device__ __forceinline__ double perbin_term(double raw_e, double m) {
    if (m > 0.0) {
        const double d = raw_e - 1e-9;
        const double es = 0.5 * (raw_e + 1e-9 + sqrt(fma(d, d, 1e-18)));
        return -(raw_e - m * log(es));
    }
    return -raw_e;
}

__global__ void __launch_bounds__(kRatesThreads)
rates_kernel(const DevTables T, const RatesLaunch L, const double* __restrict__ params,
             const double* __restrict__ osc /* [n_points/group][2][1+n_cosz][n_etrue] */, RatesOut O) {
    extern __shared__ double smem[];
    const int b = blockIdx.x;
    const int nT = T.n_etrue, nZ = T.n_cosz, nrow = 1 + nZ;
    double* S0 = smem;                       // oscillated sample, ES nu_e      [nrow][nT]
    double* S1 = S0 + (size_t)nrow * nT;     //                    ES nu_other
    double* S2 = S1 + (size_t)nrow * nT;     //                    CC
    double* bgs = S2 + (size_t)nrow * nT;    // summed backgrounds, ES then CC
    __shared__ double red[kRatesThreads / 32];
    __shared__ uint32_t redf[kRatesThreads / 32];

    const double* p = params + (size_t)b * L.param_stride;
    const double* P8 = osc + (size_t)(b / L.group) * 2 * nrow * nT;
    const double* Ph = P8 + (size_t)nrow * nT;
    const double f8 = p[3], fh = p[4];

    // earth-norm bounds (reference likelihood_debug.jl:1-4, likelihood_builder.jl:227)
    bool oob = false;
    if (T.earth_unc)
        for (int k = 0; k < T.n_layers; ++k) {
            const double x = p[SOLAROSC_N_FIXED_PARAMS + k];
            oob |= !(x >= 0.0 && x <= 2.0);
        }

    double total = 0.0;
    uint32_t fl = 0;
    for (int di = 0; di < T.n_det; ++di) {
        if (L.det_only >= 0 && di != L.det_only) continue;
        const DevDetector& D = T.det[di];
        const double* pd = p + D.param_off;
        double* bg_es = bgs;
        double* bg_cc = bgs + D.n_es;
        __syncthreads();
        // ---- backgrounds (reference propagation_bg.jl:36-83; every ES component uses norm_1) ----
        for (int j = threadIdx.x; j < D.n_es + D.n_cc; j += blockDim.x) {
            double acc = 0.0;
            if (j < D.n_es) {
                for (int c = 0; c < D.n_bg_es; ++c) {
                    const double raw = D.raw_bg_es[(size_t)c * D.n_es + j];
                    acc += (D.bg_es_area[c] != 0.0) ? raw * (D.bg_es_area[c] * pd[0]) : raw;
                }
                bg_es[j] = acc;
                if (L.mode == kModeRates && O.bg_es) O.bg_es[(size_t)b * D.n_es + j] = acc;
            } else {
                const int jc = j - D.n_es;
                for (int c = 0; c < D.n_bg_cc; ++c) {
                    const double raw = D.raw_bg_cc[(size_t)c * D.n_cc + jc];
                    acc += (D.bg_cc_norm[c] >= 0) ? raw * (pd[D.n_norm_es + D.bg_cc_norm[c]] / 2.2e-6) : raw;
                }
                bg_cc[jc] = acc;
                if (L.mode == kModeRates && O.bg_cc) O.bg_cc[(size_t)b * D.n_cc + jc] = acc;
            }
        }
        // ---- oscillated samples (reference propagation_osc.jl:159-210) ----------------------------
        for (int t = threadIdx.x; t < nrow * nT; t += blockDim.x) {
            const int row = t / nT, i = t - row * nT;
            const double p8 = P8[t], ph = Ph[t];
            if (D.es_mode) {
                if (row == 0) {
                    S0[t] = D.u_es_nue_8b[i] * p8 * f8 + D.u_es_nue_hep[i] * ph * fh;
                    S1[t] = D.u_es_nuo_8b[i] * (1.0 - p8) * f8 + D.u_es_nuo_hep[i] * (1.0 - ph) * fh;
                } else {
                    const double w = T.wz[row - 1];
                    S0[t] = D.u_es_nue_8b[i] * (f8 * p8 * w) + D.u_es_nue_hep[i] * (fh * ph * w);
                    S1[t] = D.u_es_nuo_8b[i] * (f8 * (1.0 - p8) * w) + D.u_es_nuo_hep[i] * (fh * (1.0 - ph) * w);
                }
            }
            if (D.cc_mode) {
                const double l = D.log_e_norm[i];
                const double shape = exp(fma(p[6], l, p[7] * (l * l))) * p[5];
                const double c8 = D.u_cc_8b[i] * shape, ch = D.u_cc_hep[i] * shape;
                if (row == 0) {
                    S2[t] = c8 * p8 * f8 + ch * ph * fh;
                } else {
                    const double w = T.wz[row - 1];
                    S2[t] = c8 * (f8 * p8 * w) + ch * (fh * ph * w);
                }
            }
        }
        __syncthreads();
        // ---- fold + background + Poisson: one output bin (row, j) per thread iteration -----------------
        // (reference propagation_core.jl:73-79, propagation_reco.jl:10-29,117-130,
        //  likelihood_builder.jl:5-20,129-144)
        const int n_out_es = D.es_mode ? nrow * D.n_es : 0;
        const int n_out_cc = D.cc_mode ? nrow * D.n_cc : 0;
        double part = 0.0;
        for (int o = threadIdx.x; o < n_out_es + n_out_cc; o += blockDim.x) {
            const bool is_es = o < n_out_es;
            const int oo = is_es ? o : o - n_out_es;
            const int nR = is_es ? D.n_es : D.n_cc;
            const int row = oo / nR, j = oo - row * nR;
            double rate;
            if (is_es) {
                const double* a = S0 + (size_t)row * nT;
                const double* c = S1 + (size_t)row * nT;
                double s0 = 0.0, s1 = 0.0;
                for (int i = 0; i < nT; ++i) {
                    s0 = fma(a[i], __ldg(D.r_es_nue + (size_t)i * nR + j), s0);
                    s1 = fma(c[i], __ldg(D.r_es_nuo + (size_t)i * nR + j), s1);
                }
                const double bgv = (row == 0) ? bg_es[j] : bg_es[j] * T.wz[row - 1];
                rate = (0.5 * (s0 * D.eff_es_nue[j]) + 0.5 * (s1 * D.eff_es_nuo[j])) + 0.5 * bgv;
            } else {
                const double* a = S2 + (size_t)row * nT;
                double s0 = 0.0;
                for (int i = 0; i < nT; ++i) s0 = fma(a[i], __ldg(D.r_cc + (size_t)i * nR + j), s0);
                const double bgv = (row == 0) ? bg_cc[j] : bg_cc[j] * T.wz[row - 1];
                rate = 0.5 * (s0 * D.eff_cc[j]) + 0.5 * bgv;
            }
            const int idx0 = (is_es ? D.index_es : D.index_cc) - 1;
            if (L.mode == kModeLlh) {
                if (j >= idx0) {
                    const size_t od = (size_t)row * nR + j;
                    part += poisson_term(rate, (is_es ? D.obs_es : D.obs_cc)[od], (is_es ? D.logobs_es : D.logobs_cc)[od], fl);
                }
            } else {
                double v = rate;
                if (L.mode == kModePerbin) {
                    const size_t od = (size_t)row * nR + j;
                    v = (j >= idx0 && j < nR - 1) ? perbin_term(rate, (is_es ? D.obs_es : D.obs_cc)[od]) : 0.0;
                    if (oob) v = -INFINITY;
                }
                double* day = is_es ? O.es_day : O.cc_day;
                double* night = is_es ? O.es_night : O.cc_night;
                if (row == 0) {
                    if (day) day[(size_t)b * nR + j] = v;
                } else if (night) {
                    const long long sr = is_es ? L.s_row_es : L.s_row_cc, sj = is_es ? L.s_j_es : L.s_j_cc;
                    night[(size_t)b * nZ * nR + (size_t)(row - 1) * sr + (size_t)j * sj] = v;
                }
            }
        }
        total += part;
    }
    if (L.mode != kModeLlh) return;
    // ---- block reduction (fixed order: deterministic) ----------------------------------------------
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    total = warp_sum(total);
    for (int o = 16; o > 0; o >>= 1) fl |= __shfl_xor_sync(0xffffffffu, fl, o);
    if (lane == 0) { red[warp] = total; redf[warp] = fl; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        uint32_t f = 0;
        for (int w = 0; w < kRatesThreads / 32; ++w) { s += red[w]; f |= redf[w]; }
        double lnl = -s;
        if (f & SOLAROSC_FLAG_NEG_ERROR) lnl = NAN;
        if (oob) { lnl = -INFINITY; f = SOLAROSC_FLAG_EARTH_OOB; }
        else if (lnl != lnl) f |= (f & SOLAROSC_FLAG_NEG_ERROR) ? 0u : SOLAROSC_FLAG_NONFINITE;
        O.lnl[b] = lnl;
        if (O.flags) O.flags[b] = f;
    }
}

// ------------------------------------------------------------------------------------------
// Stand-alone Poisson reduction: one CTA per (expected, measured) vector pair
// (reference poissonLogLikelihood, likelihood_statistical.jl:1-28; the shape timed by
//  tests/benchmark_likelihoods.jl:10-19)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
poisson_kernel(const double* __restrict__ e, const double* __restrict__ m, size_t len, double* __restrict__ out,
               uint32_t* __restrict__ flags) {
    __shared__ double red[8];
    __shared__ uint32_t redf[8];
    const double* ev = e + (size_t)blockIdx.x * len;
    const double* mv = m + (size_t)blockIdx.x * len;
    double acc = 0.0;
    uint32_t fl = 0;
    for (size_t i = threadIdx.x; i < len; i += blockDim.x) {
        const double mm = mv[i];
        acc += poisson_term(ev[i], mm, mm > 0.0 ? log(mm) : 0.0, fl);
    }
    acc = warp_sum(acc);
    for (int o = 16; o > 0; o >>= 1) fl |= __shfl_xor_sync(0xffffffffu, fl, o);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = acc; redf[threadIdx.x >> 5] = fl; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        uint32_t f = 0;
        for (int w = 0; w < 8; ++w) { s += red[w]; f |= redf[w]; }
        out[blockIdx.x] = (f & SOLAROSC_FLAG_NEG_ERROR) ? NAN : -s;
        if (flags) flags[blockIdx.x] = f;
    }
}

}  // namespace solarosc
