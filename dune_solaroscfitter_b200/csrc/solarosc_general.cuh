// solarosc_general.cuh -- K3 + K4 + K5 for the inclusive, semi-inclusive and angular-ES branches.
//
// Reference: propagateSamples steps 6-7 (propagation_main.jl:98-153), compute_ES_angular_event_rates,
// compute_CC_inclusive_event_rates, compute_ES_misID_CC_event_rates (propagation_reco.jl:66-108,139-199),
// llh_ES_angle (likelihood_builder.jl:53-78), use_ES / use_CC of likelihood_main.jl:8-20.
//
// Every branch is still linear in the oscillation probabilities, so it is the same constant-matrix fold on the
// FP64 tensor pipe as fold_stage, with up to two matrix pairs per output channel:
//
//   ES bins:  S1[row][j] = F8 sum_i P8 A8 + Fh sum_i Ph Ah + (F8 C8 + Fh Ch)            (ES signal, "plain" pair)
//             S2[row][j] = f_above (F8 sum_i P8 s_i X8 + Fh sum_i Ph s_i Xh) + BG_ES/2    (CC projected on ES bins, "shaped" pair)
//   CC bins:  S1[row][j] = f_below (F8 sum_i P8 s_i B8 + Fh sum_i Ph s_i Bh + BG_CC/2)
//                          + F8 sum_i P8 Y8 + Fh sum_i Ph Yh + (F8 YC8 + Fh YCh)           (mis-identified ES)
//   non-angular:  rate[row][j]    = w_row (S1 + S2)
//   angular ES:   rate[c][row][j] = w_row (mask_c ang[c][j] S1 + bgang[c][j] S2)           for every cos-scatter bin c
//
// (f_above = f_below = 1 and no mis-ID pair in the inclusive analysis, whose CC arrays are zero and stay out of the
// likelihood; X = 0.5 eff_CCincl u_cc R_CCincl, Y = 0.5 eff_cc (u_nue R_misID_nue - u_nuo R_misID_nuo).)
// Handles that contain such a detector run osc_kernel + gen_kernel for every call; the tuned exclusive kernels are
// untouched.
#pragma once
#include "solarosc_kernels.cuh"

namespace solarosc {

constexpr int kMCg = 2;   // row tiles in flight (two accumulator sets per tile)

struct GenLaunch {
    // output strides of the angular ES arrays (cos bin, reco bin, zenith bin) for day and night
    long long d_c, d_j, n_c, n_j, n_z;
};

template <int MODE>
__device__ __forceinline__ void gen_fold_vectors(const DevTables& T, const DevDetector& D, const double* __restrict__ p, size_t b,
                                                 double* scr, const RatesOut& O) {
    const int nT = T.n_etrue, kt = fold_kt(nT);
    const int nte = (D.n_es + 7) / 8, ntc = (D.n_cc + 7) / 8;
    const double f8 = p[3], fh = p[4];
    const double* pd = p + D.param_off;
    double* scale = scr;
    double* cst1_es = scr + 8 * kt;
    double* cst2_es = cst1_es + 8 * nte;
    double* cst_cc = cst2_es + 8 * nte;
    double* bgs_cc = cst_cc + 8 * ntc;     // summed CC backgrounds per reco bin (asymmetry mode)
    for (int i = threadIdx.x; i < 4 * kt; i += blockDim.x) {
        double sh = 0.0;
        if (i < nT) {
            const double l = D.log_e_norm[i];
            sh = exp(fma(p[6], l, p[7] * (l * l))) * p[5];
        }
        scale[i] = f8 * sh; scale[4 * kt + i] = fh * sh;
    }
    for (int j = threadIdx.x; j < 8 * (nte + ntc); j += blockDim.x) {
        double acc = 0.0;
        if (j < 8 * nte) {
            double c1 = 0.0;
            if (j < D.n_es) {
                for (int q = 0; q < D.n_bg_es; ++q) {
                    const double raw = D.raw_bg_es[q * D.n_es + j];
                    acc += (D.bg_es_area[q] != 0.0) ? raw * (D.bg_es_area[q] * pd[0]) : raw;
                }
                c1 = fma(f8, D.c8[j], fh * D.ch[j]);
                if (MODE == kModeRates && O.bg_es) O.bg_es[b * D.n_es + j] = acc;
            }
            cst1_es[j] = c1;
            cst2_es[j] = 0.5 * acc;
        } else {
            const int jc = j - 8 * nte;
            double c = 0.0;
            if (jc < D.n_cc) {
                for (int q = 0; q < D.n_bg_cc; ++q) {
                    const double raw = D.raw_bg_cc[q * D.n_cc + jc];
                    acc += (D.bg_cc_norm[q] >= 0) ? raw * (pd[D.n_norm_es + D.bg_cc_norm[q]] * (1.0 / 2.2e-6)) : raw;
                }
                c = D.f_below * (0.5 * acc);
                if (D.yc8) c += fma(f8, D.yc8[jc], fh * D.ych[jc]);
                if (MODE == kModeRates && O.bg_cc) O.bg_cc[b * D.n_cc + jc] = acc;
            }
            cst_cc[jc] = c;
            if (MODE == kModeAsym) bgs_cc[jc] = acc;
        }
    }
}

// K3 + K4 + K5 of one parameter point on the probability tile (p8s, phs: [rows][ldp] per process in shared memory).
// Starts with a barrier (orders the tile and the scratch), like fold_stage.
template <int MODE>
__device__ __forceinline__ void gen_fold_stage(const DevTables& T, const RatesLaunch& L, const GenLaunch& GL, const double* __restrict__ p,
                                               size_t b, const double* p8s, const double* phs, int ldp, double* scr, const RatesOut& O) {
    const int nT = T.n_etrue, nZ = T.n_cosz, nrow = 1 + nZ;
    const int kt = fold_kt(nT), mt = (nrow + 7) / 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, fg = lane >> 2, ft = lane & 3;
    double* scale = scr;
    double* red = scr + fold_vec_doubles(nT, L.max_ereco);
    uint32_t* redf = reinterpret_cast<uint32_t*>(red + kWarps);
    const double f8 = p[3], fh = p[4];
    bool oob = false;
    if (T.earth_unc)
        for (int k = 0; k < T.n_layers; ++k) {
            const double x = p[SOLAROSC_N_FIXED_PARAMS + k];
            oob |= !(x >= 0.0 && x <= 2.0);
        }
    double total = 0.0;
    uint32_t fl = 0;
    const uint32_t tile_bytes = 8u * (uint32_t)ldp * 8u;
    const uint32_t dproc = smem_addr(phs) - smem_addr(p8s);
    for (int di = 0; di < T.n_det; ++di) {
        if (L.det_only >= 0 && di != L.det_only) continue;
        const DevDetector& D = T.det[di];
        const int nte = (D.n_es + 7) / 8, ntc = (D.n_cc + 7) / 8;
        const double* cst1_es = scr + 8 * kt;
        const double* cst2_es = cst1_es + 8 * nte;
        const double* cst_cc = cst2_es + 8 * nte;
        const bool es_on = D.es_mode != 0, cc_on = D.cc_out != 0;   // channels whose arrays are computed
        const int ncos = D.ang ? D.n_cos : 1;
        __syncthreads();
        gen_fold_vectors<MODE>(T, D, p, b, scr, O);
        if (MODE == kModeRates || MODE == kModePerbin) {   // disabled channels still produce arrays (propagation_reco.jl:12-15,119-122)
            const double zv = (MODE == kModePerbin && oob) ? -INFINITY : 0.0;
            if (!es_on) {
                const size_t per_d = (size_t)ncos * D.n_es, per_n = per_d * nZ;
                for (size_t o = threadIdx.x; o < per_d + per_n; o += blockDim.x) {
                    if (o < per_d) { if (O.es_day) O.es_day[b * per_d + o] = zv; }
                    else if (O.es_night) O.es_night[b * per_n + (o - per_d)] = zv;
                }
            }
            if (!cc_on)
                for (int o = threadIdx.x; o < nrow * D.n_cc; o += blockDim.x) {
                    if (o < D.n_cc) { if (O.cc_day) O.cc_day[b * D.n_cc + o] = zv; }
                    else if (O.cc_night) O.cc_night[b * nZ * D.n_cc + (o - D.n_cc)] = zv;
                }
        }
        __syncthreads();
        double part = 0.0;
        double d_es = 0.0, n_es = 0.0, d_cc = 0.0, n_cc = 0.0;   // asymmetry mode: day / night totals above threshold
        // work items: (column block, group of kMCg row tiles), the ES blocks first (with the angular reconstruction they
        // carry n_cos times the Poisson terms), dealt round-robin over the warps
        const int npass = (mt + kMCg - 1) / kMCg;
        for (int task = warp; task < (nte + ntc) * npass; task += kWarps) {
            const bool is_es = task < nte * npass;
            const int tt = is_es ? task : task - nte * npass, nc_ = is_es ? nte : ntc;
            const int pass = tt / nc_, col = (is_es ? 0 : nte) + (tt - pass * nc_);
            if (is_es ? !es_on : !cc_on) continue;
            const int n0 = 8 * (is_es ? col : col - nte);
            const int ldb = 8 * (is_es ? nte : ntc);
            const int nR = is_es ? D.n_es : D.n_cc;
            const int idx0 = (is_es ? D.index_es : D.index_cc) - 1;
            const int j0 = n0 + 2 * ft;
            const bool v0 = j0 < nR, v1 = j0 + 1 < nR;
            const size_t boff = (size_t)ft * ldb + n0 + fg;
            // plain pair -> S1, shaped pair -> S2 (ES) or S1 (CC)
            const double *pl8 = is_es ? D.a8 : D.yp8, *plh = is_es ? D.ah : D.yph;
            const double *sh8 = is_es ? D.xs8 : D.b8, *shh = is_es ? D.xsh : D.bh;
            const uint32_t sc8 = smem_addr(scale + ft), sch = sc8 + 4u * kt * 8u;
            const double i10 = v0 ? (is_es ? cst1_es : cst_cc)[j0] : 0.0, i11 = v1 ? (is_es ? cst1_es : cst_cc)[j0 + 1] : 0.0;
            const double i20 = (v0 && is_es) ? cst2_es[j0] : 0.0, i21 = (v1 && is_es) ? cst2_es[j0 + 1] : 0.0;
            const size_t bstep = (size_t)4 * ldb;
            {
                const int m0 = pass * kMCg;
                double a1[kMCg][2], a2[kMCg][2];
                uint32_t pa[kMCg];
#pragma unroll
                for (int c = 0; c < kMCg; ++c) {
                    a1[c][0] = i10; a1[c][1] = i11; a2[c][0] = i20; a2[c][1] = i21;
                    pa[c] = smem_addr(p8s + fg * ldp + ft) + (uint32_t)min(m0 + c, mt - 1) * tile_bytes;
                }
                for (int ks = 0; ks < kt; ++ks) {
                    if (pl8) {
                        const double b8v = __ldg(pl8 + boff + ks * bstep) * f8, bhv = __ldg(plh + boff + ks * bstep) * fh;
#pragma unroll
                        for (int c = 0; c < kMCg; ++c) {
                            dmma_8x8x4(a1[c][0], a1[c][1], lds_f64(pa[c] + 32u * ks), b8v);
                            dmma_8x8x4(a1[c][0], a1[c][1], lds_f64(pa[c] + dproc + 32u * ks), bhv);
                        }
                    }
                    if (sh8) {
                        const double b8v = __ldg(sh8 + boff + ks * bstep) * lds_f64(sc8 + 32u * ks);
                        const double bhv = __ldg(shh + boff + ks * bstep) * lds_f64(sch + 32u * ks);
#pragma unroll
                        for (int c = 0; c < kMCg; ++c) {
                            if (is_es) {
                                dmma_8x8x4(a2[c][0], a2[c][1], lds_f64(pa[c] + 32u * ks), b8v);
                                dmma_8x8x4(a2[c][0], a2[c][1], lds_f64(pa[c] + dproc + 32u * ks), bhv);
                            } else {
                                dmma_8x8x4(a1[c][0], a1[c][1], lds_f64(pa[c] + 32u * ks), b8v);
                                dmma_8x8x4(a1[c][0], a1[c][1], lds_f64(pa[c] + dproc + 32u * ks), bhv);
                            }
                        }
                    }
                }
                // ---- epilogue ----------------------------------------------------------------------------
                double r1[kMCg][2], r2[kMCg][2];
                bool vr[kMCg];
#pragma unroll
                for (int c = 0; c < kMCg; ++c) {
                    const int row = 8 * (m0 + c) + fg;
                    vr[c] = m0 + c < mt && row < nrow;
                    const double w = (row == 0) ? 1.0 : T.wz[min(row, nrow - 1) - 1];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) { r1[c][hh] = w * a1[c][hh]; r2[c][hh] = w * a2[c][hh]; }
                }
                if (!(is_es && D.ang)) {
#pragma unroll
                    for (int c = 0; c < kMCg; ++c) {
                        const int row = 8 * (m0 + c) + fg;
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const int jx = j0 + hh;
                            if (!vr[c] || jx >= nR) continue;
                            const double r = r1[c][hh] + r2[c][hh];
                            if (MODE == kModeLlh) {
                                if (jx >= idx0) {
                                    const double2 ol = __ldg((is_es ? D.ol_es : D.ol_cc) + row * nR + jx);
                                    part += poisson_term(r, ol.x, ol.y, fl);
                                }
                            } else if (MODE == kModeAsym) {
                                if (jx >= idx0) {
                                    if (is_es) { if (row == 0) d_es += r; else n_es += r; }
                                    else { if (row == 0) d_cc += r; else n_cc += r; }
                                }
                            } else {
                                double v = r;
                                if (MODE == kModePerbin) {   // likelihood_builder.jl:22-51,146-174: bins idx..n-1, the overflow bin excluded
                                    v = (jx >= idx0 && jx < nR - 1) ? perbin_term(r, __ldg((is_es ? D.ol_es : D.ol_cc) + row * nR + jx).x) : 0.0;
                                    if (oob) v = -INFINITY;
                                }
                                double* day = is_es ? O.es_day : O.cc_day;
                                double* night = is_es ? O.es_night : O.cc_night;
                                const long long sr = is_es ? L.s_row_es : L.s_row_cc, sj = is_es ? L.s_j_es : L.s_j_cc;
                                if (row == 0) { if (day) day[b * nR + jx] = v; }
                                else if (night) night[b * nZ * nR + (size_t)(row - 1) * sr + (size_t)jx * sj] = v;
                            }
                        }
                    }
                } else {
                    // every cos-scatter bin of this lane's (row, reco bin) elements: llh_ES_angle sums them all.  The cos
                    // bin is the outer loop so that the 2 kMCg Poisson terms of one bin are independent chains.
                    const bool in0 = v0 && j0 >= idx0, in1 = v1 && j0 + 1 >= idx0;
                    const double* am0 = D.angm + (size_t)min(j0, nR - 1) * ncos;
                    const double* am1 = D.angm + (size_t)min(j0 + 1, nR - 1) * ncos;
                    const double* bm0 = D.bgang + (size_t)min(j0, nR - 1) * ncos;
                    const double* bm1 = D.bgang + (size_t)min(j0 + 1, nR - 1) * ncos;
                    if (MODE == kModeAsym) {
                        // the asymmetry sums every cos-scatter bin (derive_variables_from_chain.jl:239-241): the column sums
                        // of the angular tables do that once and for all
#pragma unroll
                        for (int c = 0; c < kMCg; ++c) {
                            if (!vr[c]) continue;
                            double s01 = 0.0;
                            if (in0) s01 += fma(D.ang_sa[j0], r1[c][0], D.ang_sb[j0] * r2[c][0]);
                            if (in1) s01 += fma(D.ang_sa[j0 + 1], r1[c][1], D.ang_sb[j0 + 1] * r2[c][1]);
                            if (8 * (m0 + c) + fg == 0) d_es += s01; else n_es += s01;
                        }
                    } else
                    for (int ci = 0; ci < (MODE == kModeLlh ? D.n_cos_live : ncos); ++ci) {
                        const int cs = MODE == kModeLlh ? (int)D.cos_live[ci] : ci;   // the likelihood skips the bins the cut empties
                        const double ga0 = __ldg(am0 + cs), ga1 = __ldg(am1 + cs), gb0 = __ldg(bm0 + cs), gb1 = __ldg(bm1 + cs);
                        if (MODE == kModeLlh) {
                            double2 o0[kMCg], o1[kMCg];
#pragma unroll
                            for (int c = 0; c < kMCg; ++c) {
                                const int row = min(8 * (m0 + c) + fg, nrow - 1);
                                o0[c] = (vr[c] && in0) ? __ldg(D.ol_es_ang + ((size_t)row * nR + j0) * ncos + cs) : make_double2(0.0, 0.0);
                                o1[c] = (vr[c] && in1) ? __ldg(D.ol_es_ang + ((size_t)row * nR + j0 + 1) * ncos + cs) : make_double2(0.0, 0.0);
                            }
#pragma unroll
                            for (int c = 0; c < kMCg; ++c) {
                                const double e0 = fma(ga0, r1[c][0], gb0 * r2[c][0]), e1 = fma(ga1, r1[c][1], gb1 * r2[c][1]);
                                const double t0 = poisson_term((vr[c] && in0) ? e0 : 1.0, o0[c].x, o0[c].y, fl);
                                const double t1 = poisson_term((vr[c] && in1) ? e1 : 1.0, o1[c].x, o1[c].y, fl);
                                part += ((vr[c] && in0) ? t0 : 0.0) + ((vr[c] && in1) ? t1 : 0.0);
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < kMCg; ++c) {
                                const int row = 8 * (m0 + c) + fg;
                                if (!vr[c]) continue;
#pragma unroll
                                for (int hh = 0; hh < 2; ++hh) {
                                    const int jx = j0 + hh;
                                    if (jx >= nR) continue;
                                    double r = fma(hh ? ga1 : ga0, r1[c][hh], (hh ? gb1 : gb0) * r2[c][hh]);
                                    if (MODE == kModePerbin) {   // llh_ES_angle_perbin, likelihood_builder.jl:80-108
                                        r = (jx >= idx0 && jx < nR - 1) ? perbin_term(r, __ldg(D.ol_es_ang + ((size_t)row * nR + jx) * ncos + cs).x) : 0.0;
                                        if (oob) r = -INFINITY;
                                    }
                                    if (row == 0) {
                                        if (O.es_day) O.es_day[b * (size_t)ncos * nR + cs * GL.d_c + jx * GL.d_j] = r;
                                    } else if (O.es_night) {
                                        O.es_night[b * (size_t)ncos * nR * nZ + cs * GL.n_c + jx * GL.n_j + (row - 1) * GL.n_z] = r;
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        if (MODE == kModeAsym) {
            // derive_variables_from_chain.jl:236-258: half of the background above threshold comes off the day total and half
            // off the night total; a zero denominator gives 0; the CC asymmetry is not formed in the inclusive analysis
            for (int ch = 0; ch < 2; ++ch) {
                const bool is_es = ch == 0;
                double* out = is_es ? O.asym_es : O.asym_cc;
                if (is_es ? !es_on : !cc_on) {
                    if (out && threadIdx.x == 0) out[b] = 0.0;
                    continue;
                }
                const double dtot = block_sum(is_es ? d_es : d_cc, red), ntot = block_sum(is_es ? n_es : n_cc, red);
                if (threadIdx.x == 0) {
                    const int nR = is_es ? D.n_es : D.n_cc, idx0 = (is_es ? D.index_es : D.index_cc) - 1;
                    double bg = 0.0;
                    for (int j = idx0; j < nR; ++j) bg += is_es ? 2.0 * cst2_es[j] : (cst_cc + 8 * ntc)[j];
                    const double dd = dtot - 0.5 * bg, nn = ntot - 0.5 * bg, den = dd + nn;
                    if (out) out[b] = (den == 0.0) ? 0.0 : 2.0 * (dd - nn) / den;
                }
            }
        }
        if (MODE == kModeLlh && D.ang && es_on && threadIdx.x == 0) part += D.llh_dead_const;
        total += part;
    }
    if (MODE != kModeLlh) return;
    total = warp_sum(total);
    for (int o = 16; o > 0; o >>= 1) fl |= __shfl_xor_sync(0xffffffffu, fl, o);
    __syncthreads();
    if (lane == 0) { red[warp] = total; redf[warp] = fl; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        uint32_t f = 0;
        for (int w = 0; w < kWarps; ++w) { s += red[w]; f |= redf[w]; }
        double lnl = -s;
        if (f & SOLAROSC_FLAG_NEG_ERROR) lnl = NAN;
        if (oob) { lnl = -INFINITY; f = SOLAROSC_FLAG_EARTH_OOB; }
        else if (lnl != lnl) f |= (f & SOLAROSC_FLAG_NEG_ERROR) ? 0u : SOLAROSC_FLAG_NONFINITE;
        O.lnl[b] = lnl;
        if (O.flags) O.flags[b] = f;
    }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, SOLAROSC_RATES_MINBLOCKS)
gen_kernel(const DevTables T, const RatesLaunch L, const GenLaunch GL, const double* __restrict__ params,
           const double* __restrict__ osc /* [n_points/group][2][1+n_cosz][n_etrue] */, RatesOut O) {
    extern __shared__ __align__(16) double smem[];
    const int nT = T.n_etrue, nrow = 1 + T.n_cosz, ldp = fold_ldp(nT), rows = fold_rows(nrow);
    double* p8s = smem;
    double* phs = p8s + (size_t)rows * ldp;
    double* scr = p8s + fold_p_doubles(nT, nrow);
    const size_t b = blockIdx.x;
    const double* src = osc + (b / L.group) * 2 * nrow * nT;
    fold_zero_padding(p8s, nT, rows, ldp);
    {
        const int dq = (int)blockDim.x / nT, dr = (int)blockDim.x - dq * nT;
        int row = (int)threadIdx.x / nT, i = (int)threadIdx.x - row * nT;
        while (row < nrow) {
            p8s[row * ldp + i] = src[row * nT + i];
            phs[row * ldp + i] = src[(nrow + row) * nT + i];
            i += dr; row += dq;
            if (i >= nT) { i -= nT; ++row; }
        }
    }
    gen_fold_stage<MODE>(T, L, GL, params + b * L.param_stride, b, p8s, phs, ldp, scr, O);
}

// Fused variant (likelihood and asymmetry): K1 + K2 into the shared-memory probability tile as in llh_kernel, then the
// general fold for every point of the group.  Same shared-memory layout as llh_kernel.
template <int MODE>
__global__ void __launch_bounds__(kThreads, SOLAROSC_LLH_MINBLOCKS)
gen_llh_kernel(const DevTables T, const RatesLaunch L, const GenLaunch GL, const double* __restrict__ params, RatesOut O) {
    extern __shared__ __align__(16) double smem[];
    const int nT = T.n_etrue, nrow = 1 + T.n_cosz, ldp = fold_ldp(nT);
    const int cbs = L.osc_cb;
    double* tbl = smem;
    double* msw = smem + (size_t)kEntry * T.n_ltbl * 2 * cbs;
    double* p8s = smem + osc_smem_doubles(T.n_ltbl, cbs);
    double* phs = p8s + (size_t)nrow * ldp;
    double* scr = phs + (size_t)nrow * ldp;
    const SegTables G = stage_seg_tables(T, scr + fold_scratch_doubles(nT, L.max_ereco), true);
    fold_zero_padding(p8s, nT, nrow, ldp);
    const int n_osc = L.n_points / L.group;
    for (int g = blockIdx.x; g < n_osc; g += gridDim.x) {
        const double* po = params + (size_t)g * L.group * L.param_stride;
        for (int i0 = 0; i0 < nT; i0 += cbs) {
            if (i0 > 0 || g != (int)blockIdx.x) __syncthreads();   // the previous table / tile are no longer being read
            osc_stage<true>(T, G, po, i0, min(cbs, nT - i0), cbs, tbl, msw, nullptr, SmemSink{p8s, phs, ldp});
        }
        for (int q = 0; q < L.group; ++q) {
            const size_t b = (size_t)g * L.group + q;
            gen_fold_stage<MODE>(T, L, GL, params + b * L.param_stride, b, p8s, phs, ldp, scr, O);
        }
    }
}

// CC_incl_spectrum (the 7th element of propagateSamples' tuple, propagation_main.jl:108-146; setup.jl:235-243 prints the
// split of the inclusive data with it): the inclusive-CC part of the ES arrays summed over day + night rows (and over the
// cos-scatter bins of the angular reconstruction),
//   spectrum[j] = sa_j sum_i s_i (F8 X8[i][j] Q8[i] + Fh Xh[i][j] Qh[i]),   Q[i] = sum_rows w_row P[row][i],
// X = f_above 0.5 eff_CCincl u_cc R_CCincl (D.xs8 / D.xsh), sa_j = sum_c BG.angular[c][j] (1 when not angular).
// One CTA per point on the probabilities osc_kernel wrote.
__global__ void __launch_bounds__(kThreads)
cc_incl_spectrum_kernel(const DevTables T, int det, int param_stride, const double* __restrict__ params,
                        const double* __restrict__ osc /* [n][2][1+n_cosz][n_etrue] */, double* __restrict__ out /* [n][n_es] */) {
    extern __shared__ __align__(16) double smem[];
    const DevDetector& D = T.det[det];
    const int nT = T.n_etrue, nrow = 1 + T.n_cosz, le = 8 * ((D.n_es + 7) / 8);
    const size_t b = blockIdx.x;
    const double* p = params + b * param_stride;
    const double* p8 = osc + b * 2 * nrow * nT;
    const double* ph = p8 + (size_t)nrow * nT;
    double* q8 = smem;          // [nT] F8 s_i Q8[i]
    double* qh = smem + nT;     // [nT] Fh s_i Qh[i]
    for (int i = threadIdx.x; i < nT; i += blockDim.x) {
        double a8 = p8[i], ah = ph[i];                                   // day row, w = 1
        for (int row = 1; row < nrow; ++row) {
            const double w = T.wz[row - 1];
            a8 = fma(w, p8[(size_t)row * nT + i], a8);
            ah = fma(w, ph[(size_t)row * nT + i], ah);
        }
        const double l = D.log_e_norm[i];
        const double sh = exp(fma(p[6], l, p[7] * (l * l))) * p[5];       // cross-section shape (propagation_osc.jl:185-196)
        q8[i] = p[3] * sh * a8;
        qh[i] = p[4] * sh * ah;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < D.n_es; j += blockDim.x) {
        double acc = 0.0;
        if (D.xs8)
            for (int i = 0; i < nT; ++i) acc = fma(q8[i], D.xs8[(size_t)i * le + j], fma(qh[i], D.xsh[(size_t)i * le + j], acc));
        out[b * D.n_es + j] = D.ang ? acc * D.ang_sb[j] : acc;
    }
}

}  // namespace solarosc
