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

#ifndef SOLAROSC_OSC_THREADS
#define SOLAROSC_OSC_THREADS 256
#endif
#ifndef SOLAROSC_OSC_MINBLOCKS
#define SOLAROSC_OSC_MINBLOCKS 2
#endif
constexpr int kOscThreads = SOLAROSC_OSC_THREADS;
constexpr int kRatesThreads = 256;
constexpr int kRT = 7;   // rows of the fold kept in registers per thread

struct OscLaunch {
    int n_points;     // oscillation points in this launch
    int n_chunks;     // energy chunks per point
    int cb;           // coarse energy bins per chunk
    int param_stride; // doubles between consecutive points' parameter rows
    int point_step;   // parameter row of oscillation point g is g*point_step (grouped evaluation)
};

__host__ __device__ inline size_t osc_smem_bytes(int n_layers, int cb) {
    return sizeof(double) * ((size_t)kEntry * n_layers * 2 * cb + (size_t)4 * 2 * cb);
}

// ------------------------------------------------------------------------------------------
// K1 + K2
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kOscThreads, SOLAROSC_OSC_MINBLOCKS)
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
    double* tbl = smem;                                  // [nl][2 parity][cbs] entries of kEntry doubles
    double* msw = smem + (size_t)kEntry * nl * 2 * cbs;  // [2 proc][2 what][2 parity][cbs]
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
        const int slot = (el & 1) * cbs + (el >> 1);
        make_entry(dm21, a, e, u1, u2, u3, tbl + ((size_t)k * 2 * cbs + slot) * kEntry);
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
        const double* t0 = tbl + (size_t)j * kEntry;           // even fine energy of this coarse bin, layer 0
        const double* t1 = t0 + (size_t)cbs * kEntry;          // odd fine energy
        const int ls = 2 * cbs * kEntry;                        // doubles between consecutive layers
        for (int sub = 0; sub < T.n_sub; ++sub) {
            const int path = z * T.n_sub + sub;
            const int s0 = T.path_start[path], s1 = T.path_start[path + 1];
            // Both fine energies of the bin advance through the layers together: two independent
            // dependency chains per thread keep the FP64 pipe fed (osc_reduce_from_lookup, osc.jl:436-446).
            State pa, pb;
            {
                const int k = T.app_layer[s0] * ls;
                const double len = T.app_len[s0];
                if (s1 - s0 == 1) {
                    double y1r, y1i, y2r, y2i;
                    step_phases(t0 + k, len, y1r, y1i, y2r, y2i);
                    pa.ar = fma(t0[k + 3], y2r * t0[k + 9], fma(t0[k], y1r * t0[k + 8], u1));
                    pa.ai = fma(t0[k + 3], y2i * t0[k + 9], t0[k] * (y1i * t0[k + 8]));
                    step_phases(t1 + k, len, y1r, y1i, y2r, y2i);
                    pb.ar = fma(t1[k + 3], y2r * t1[k + 9], fma(t1[k], y1r * t1[k + 8], u1));
                    pb.ai = fma(t1[k + 3], y2i * t1[k + 9], t1[k] * (y1i * t1[k + 8]));
                } else {
                    step_first(t0 + k, len, u1, u2, u3, pa);
                    step_first(t1 + k, len, u1, u2, u3, pb);
                    for (int s = s0 + 1; s < s1 - 1; ++s) {
                        const int km = T.app_layer[s] * ls;
                        const double lm = T.app_len[s];
                        step_mid(t0 + km, lm, pa);
                        step_mid(t1 + km, lm, pb);
                    }
                    const int kl = T.app_layer[s1 - 1] * ls;
                    const double ll = T.app_len[s1 - 1];
                    step_last(t0 + kl, ll, pa);
                    step_last(t1 + kl, ll, pb);
                }
            }
            const double pe0 = fma(pa.ar, pa.ar, pa.ai * pa.ai);   // |<nu_1|X|nu_e>|^2 (osc.jl:444)
            const double pe1 = fma(pb.ar, pb.ar, pb.ai * pb.ai);
            if (p1e_out) {
                double* o = p1e_out + ((size_t)g * T.n_paths + path) * T.n_e + 2 * (i0 + j);
                o[0] = pe0; o[1] = pe1;
            }
            // IYSW night probability (osc.jl:204): P_day + c13^2 cos2theta_m (P1e - P0)
            const double d0 = pe0 - p0, d1 = pe1 - p0;
            acc8 += fma(msw[(size_t)1 * 2 * cbs + j], d0, msw[j]) + fma(msw[(size_t)1 * 2 * cbs + cbs + j], d1, msw[cbs + j]);
            acch += fma(msw[(size_t)3 * 2 * cbs + j], d0, msw[(size_t)2 * 2 * cbs + j]) +
                    fma(msw[(size_t)3 * 2 * cbs + cbs + j], d1, msw[(size_t)2 * 2 * cbs + cbs + j]);
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
    int max_ereco;     // largest reco-bin count of any detector/channel (sizes the shared-memory tiles)
    long long s_row_es, s_j_es, s_row_cc, s_j_cc;  // output strides of the night arrays (layout)
};

struct RatesOut {
    double* lnl; uint32_t* flags;
    double *es_day, *es_night, *cc_day, *cc_night, *bg_es, *bg_cc;
};

// rows of the oscillated sample are padded to a multiple of the per-pass row count and then to
// pitch = 1 (mod 8) so that the transposing stores of consecutive energies hit distinct banks
__host__ __device__ inline int rates_row_pitch(int nrow) {
    const int per_pass = (kRatesThreads / 32) * kRT;
    int p = ((nrow + per_pass - 1) / per_pass) * per_pass;
    while (p % 8 != 1) ++p;
    return p;
}

__host__ __device__ inline size_t rates_smem_bytes(int n_etrue, int n_cosz, int max_ereco) {
    const size_t ld = (size_t)((n_etrue + 1) & ~1);
    return sizeof(double) * ((size_t)3 * rates_row_pitch(1 + n_cosz) * ld + 2 * ld + (size_t)2 * max_ereco +
                             (size_t)2 * (1 + n_cosz) * max_ereco + 16);
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

template <int MODE>
__global__ void __launch_bounds__(kRatesThreads)
rates_kernel(const DevTables T, const RatesLaunch L, const double* __restrict__ params,
             const double* __restrict__ osc /* [n_points/group][2][1+n_cosz][n_etrue] */, RatesOut O) {
    extern __shared__ double smem[];
    const int b = blockIdx.x;
    const int nT = T.n_etrue, nZ = T.n_cosz, nrow = 1 + nZ;
    const int ldS = (nT + 1) & ~1;           // true energies padded to an even count
    const int npair = ldS >> 1;
    const int nrp = rates_row_pitch(nrow);   // padded row count of the pair-major oscillated sample
    double* S0 = smem;                       // oscillated sample, ES nu_e      [npair][nrp] double2
    double* S1 = S0 + (size_t)nrp * ldS;     //                    ES nu_other
    double* S2 = S1 + (size_t)nrp * ldS;     //                    CC
    double* ccs = S2 + (size_t)nrp * ldS;    // CC unoscillated sample x cross-section shape, 8B then hep
    double* bgs = ccs + 2 * (size_t)ldS;     // summed backgrounds, ES then CC
    double* rates_es = bgs + 2 * (size_t)L.max_ereco;            // folded rates [nrow][n_ereco]
    double* rates_cc = rates_es + (size_t)nrow * L.max_ereco;
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
                if (MODE == kModeRates && O.bg_es) O.bg_es[(size_t)b * D.n_es + j] = acc;
            } else {
                const int jc = j - D.n_es;
                for (int c = 0; c < D.n_bg_cc; ++c) {
                    const double raw = D.raw_bg_cc[(size_t)c * D.n_cc + jc];
                    acc += (D.bg_cc_norm[c] >= 0) ? raw * (pd[D.n_norm_es + D.bg_cc_norm[c]] / 2.2e-6) : raw;
                }
                bg_cc[jc] = acc;
                if (MODE == kModeRates && O.bg_cc) O.bg_cc[(size_t)b * D.n_cc + jc] = acc;
            }
        }
        // ---- CC cross-section shape, once per true-energy bin (reference propagation_osc.jl:185-196) ----
        for (int i = threadIdx.x; i < nT; i += blockDim.x) {
            const double l = D.log_e_norm[i];
            const double shape = exp(fma(p[6], l, p[7] * (l * l))) * p[5];
            ccs[i] = D.u_cc_8b[i] * shape;
            ccs[nT + i] = D.u_cc_hep[i] * shape;
        }
        __syncthreads();
        // ---- oscillated samples (reference propagation_osc.jl:159-210) ----------------------------
        // Stored pair-major, S[(i/2)][row] as double2 = (S[row][i], S[row][i+1]): the fold below reads
        // the rows r, r+8, r+16, ... of one energy pair at constant offsets from a single pointer.
        // (row, i) walk all nrow x nT elements with a stride of blockDim.x, updated incrementally so the
        // loop carries no integer division.
        if (nT & 1)   // odd number of true-energy bins: the pad element of the last pair must read as 0
            for (int row = threadIdx.x; row < nrow; row += blockDim.x) {
                const int dst = 2 * ((npair - 1) * nrp + row) + 1;
                S0[dst] = 0.0; S1[dst] = 0.0; S2[dst] = 0.0;
            }
        {
            const int dq = (int)blockDim.x / nT, dr = (int)blockDim.x - dq * nT;
            int row = (int)threadIdx.x / nT, i = (int)threadIdx.x - row * nT;
            while (row < nrow) {
                const int dst = 2 * ((i >> 1) * nrp + row) + (i & 1);
                const double p8 = P8[row * nT + i], ph = Ph[row * nT + i];
                const double w = (row == 0) ? 1.0 : T.wz[row - 1];
                if (D.es_mode) {
                    if (row == 0) {
                        S0[dst] = D.u_es_nue_8b[i] * p8 * f8 + D.u_es_nue_hep[i] * ph * fh;
                        S1[dst] = D.u_es_nuo_8b[i] * (1.0 - p8) * f8 + D.u_es_nuo_hep[i] * (1.0 - ph) * fh;
                    } else {
                        S0[dst] = D.u_es_nue_8b[i] * (f8 * p8 * w) + D.u_es_nue_hep[i] * (fh * ph * w);
                        S1[dst] = D.u_es_nuo_8b[i] * (f8 * (1.0 - p8) * w) + D.u_es_nuo_hep[i] * (fh * (1.0 - ph) * w);
                    }
                }
                if (D.cc_mode) {
                    const double c8 = ccs[i], ch = ccs[nT + i];
                    S2[dst] = (row == 0) ? c8 * p8 * f8 + ch * ph * fh : c8 * (f8 * p8 * w) + ch * (fh * ph * w);
                }
                i += dr; row += dq;
                if (i >= nT) { i -= nT; ++row; }
            }
        }
        __syncthreads();
        // ---- fold + background + Poisson (reference propagation_core.jl:73-79,
        //      propagation_reco.jl:10-29,117-130, likelihood_builder.jl:5-20,129-144) ------------------
        // Lane = reconstructed-energy bin j, warp w owns rows w, w+8, ...: each response element
        // R[i][j] is loaded once per thread and reused for kRT rows; the oscillated sample comes
        // from shared memory as warp-wide broadcasts of (i, i+1) pairs.
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        constexpr int kW = kRatesThreads / 32;
        for (int ch = 0; ch < 2; ++ch) {
            const bool is_es = ch == 0;
            if (is_es ? !D.es_mode : !D.cc_mode) continue;
            const int nR = is_es ? D.n_es : D.n_cc;
            const double* bgv = is_es ? bg_es : bg_cc;
            double* rout = is_es ? rates_es : rates_cc;
            for (int j0 = 0; j0 < nR; j0 += 32) {
                const int j = min(j0 + lane, nR - 1);
                const bool jv = j0 + lane < nR;
                const double effa = is_es ? D.eff_es_nue[j] : D.eff_cc[j];
                const double effb = is_es ? D.eff_es_nuo[j] : 0.0;
                const double bgj = bgv[j];
                for (int r0 = warp; r0 < nrow; r0 += kW * kRT) {
                    double acca[kRT], accb[kRT];
#pragma unroll
                    for (int k = 0; k < kRT; ++k) { acca[k] = 0.0; accb[k] = 0.0; }
                    if (is_es) {
                        const double2* sa = reinterpret_cast<const double2*>(S0) + r0;
                        const double2* sb = reinterpret_cast<const double2*>(S1) + r0;
                        const double* ra = D.r_es_nue + j;
                        const double* rb = D.r_es_nuo + j;
                        for (int i2 = 0; i2 < npair; ++i2) {
                            const double ra0 = __ldg(ra), ra1 = __ldg(ra + nR);
                            const double rb0 = __ldg(rb), rb1 = __ldg(rb + nR);
#pragma unroll
                            for (int k = 0; k < kRT; ++k) {
                                const double2 va = sa[k * kW], vb = sb[k * kW];
                                acca[k] = fma(va.y, ra1, fma(va.x, ra0, acca[k]));
                                accb[k] = fma(vb.y, rb1, fma(vb.x, rb0, accb[k]));
                            }
                            sa += nrp; sb += nrp; ra += 2 * nR; rb += 2 * nR;
                        }
                    } else {
                        const double2* sa = reinterpret_cast<const double2*>(S2) + r0;
                        const double* ra = D.r_cc + j;
                        for (int i2 = 0; i2 < npair; ++i2) {
                            const double ra0 = __ldg(ra), ra1 = __ldg(ra + nR);
#pragma unroll
                            for (int k = 0; k < kRT; ++k) {
                                const double2 va = sa[k * kW];
                                acca[k] = fma(va.y, ra1, fma(va.x, ra0, acca[k]));
                            }
                            sa += nrp; ra += 2 * nR;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < kRT; ++k) {
                        const int row = r0 + k * kW;
                        if (row < nrow && jv) {
                            const double bgr = (row == 0) ? bgj : bgj * T.wz[row - 1];
                            rout[row * nR + j] = is_es ? (0.5 * (acca[k] * effa) + 0.5 * (accb[k] * effb)) + 0.5 * bgr
                                                       : 0.5 * (acca[k] * effa) + 0.5 * bgr;
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- likelihood terms / outputs, one reconstructed bin per thread iteration ------------------
        double part = 0.0;
        for (int ch = 0; ch < 2; ++ch) {
            const bool is_es = ch == 0;
            if (is_es ? !D.es_mode : !D.cc_mode) continue;
            const int nR = is_es ? D.n_es : D.n_cc;
            const double* rin = is_es ? rates_es : rates_cc;
            const double* obs = is_es ? D.obs_es : D.obs_cc;
            const double* lobs = is_es ? D.logobs_es : D.logobs_cc;
            const int idx0 = (is_es ? D.index_es : D.index_cc) - 1;
            double* day = is_es ? O.es_day : O.cc_day;
            double* night = is_es ? O.es_night : O.cc_night;
            const long long sr = is_es ? L.s_row_es : L.s_row_cc, sj = is_es ? L.s_j_es : L.s_j_cc;
            // likelihood mode only visits the bins above threshold: nc columns from idx0
            const int jlo = (MODE == kModeLlh) ? idx0 : 0, nc = nR - jlo;
            const int dq = (int)blockDim.x / nc, dr = (int)blockDim.x - dq * nc;
            int row = (int)threadIdx.x / nc, jj = (int)threadIdx.x - row * nc;
            while (row < nrow) {
                const int j = jlo + jj, od = row * nR + j;
                const double rate = rin[od];
                if (MODE == kModeLlh) {
                    part += poisson_term(rate, obs[od], lobs[od], fl);
                } else {
                    double v = rate;
                    if (MODE == kModePerbin) {
                        v = (j >= idx0 && j < nR - 1) ? perbin_term(rate, obs[od]) : 0.0;
                        if (oob) v = -INFINITY;
                    }
                    if (row == 0) {
                        if (day) day[(size_t)b * nR + j] = v;
                    } else if (night) {
                        night[(size_t)b * nZ * nR + (size_t)(row - 1) * sr + (size_t)j * sj] = v;
                    }
                }
                jj += dr; row += dq;
                if (jj >= nc) { jj -= nc; ++row; }
            }
        }
        total += part;
    }
    if (MODE != kModeLlh) return;
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

// ------------------------------------------------------------------------------------------
// FP64 FMA peak probe: 16 independent dependent-FMA chains per thread, nothing else in the loop.
// This is the denominator of the FP64 roofline (MEASURED_PEAKS.json has no FP64 figure).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double m) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 1.0 + 1e-3 * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, 1e-9);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 12345.678) out[threadIdx.x] = s;   // never true: keeps the chains alive
}

}  // namespace solarosc
