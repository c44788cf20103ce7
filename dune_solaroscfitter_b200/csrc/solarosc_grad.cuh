// solarosc_grad.cuh -- value and gradient of the log-likelihood for gradient-based samplers.
//
// The reference's MCMC mode hands BAT a log-density that is differentiated with ForwardDiff (mcmc.jl:102-103,
// likelihood_main.jl:30-39): one dual-number pass per parameter chunk through propagateSamples and the Poisson
// sum.  Here the gradient is assembled in reverse from the structure of the path:
//
//   lnL = - sum_{rows, bins} term(r),   r[row][j] = w_row x[row][j],
//   x_ES = F8 (P8 A8 + C8) + Fh (Ph Ah + Ch) + BG_ES/2,      x_CC = F8 P8 diag(s) B8 + Fh Ph diag(s) Bh + BG_CC/2.
//
//   G[row][j] = d lnL / d x[row][j] comes out of the forward fold's epilogue (analytic derivative of the
//   soft-floored Poisson term).  Fluxes, background norms and the three cross-section parameters are linear or
//   diagonal in x, so their derivatives are reductions of G against pieces the forward fold already holds.
//   The BACKWARD fold  Pbar[row][i] = d lnL / d P[row][i] = F (G_ES A^T + diag(s) G_CC B^T)  runs on the FP64
//   tensor pipe like the forward one.  The oscillation parameters (sin2_th12, sin2_th13, dm2_21, earth_norm_k)
//   enter only through the probabilities: d lnL / d theta = <Pbar, dP/d theta>.  grad_kernel writes Pbar and
//   adjoint_kernel (solarosc_adjoint.cuh) contracts it with the exact derivatives of the oscillation stage in one
//   backward sweep per (path, energy) -- every column of the gradient is exact to rounding, like the reference's
//   dual numbers.
#pragma once
#include "solarosc_kernels.cuh"

namespace solarosc {

constexpr int kGradShared = 6;                     // lnL, F8, Fh, xsec norm/tilt/curv
constexpr int kGradBg = 1 + SOLAROSC_MAX_BG;       // ES norm_1, CC norms

struct GradLaunch {
    int n_points, param_stride, max_ereco;
};

// leading dimension of a shared-memory A operand of the given K extent: multiple of 4, congruent 12 mod 16
__host__ __device__ inline int fold_ld(int width) {
    int l = 4 * ((width + 3) / 4);
    while (l % 16 != 12) l += 4;
    return l;
}
__host__ __device__ inline size_t grad_scratch_doubles(int n_etrue, int max_ereco) {
    const size_t kt4 = 4 * fold_kt(n_etrue), nt8 = 8 * ((max_ereco + 7) / 8);
    return 4 * kt4 + 4 * nt8 + (size_t)kWarps * (kGradShared > kGradBg ? kGradShared : kGradBg) + kGradShared + kGradBg + 8;
}
// Three tile pairs (P, Pbar, G) of exactly 1 + n_cosz rows each -- the last 8-row MMA tile of a pair reads past its rows into
// whatever follows, which only feeds output rows the epilogues drop (an MMA output row depends on the same A row only) --
// so that two CTAs fit one SM (111 kB each in the default binning).  `tail` pads the odd case where what follows the last
// tile is shorter than that overrun.
__host__ __device__ inline size_t grad_smem_bytes(int n_etrue, int n_cosz, int max_ereco) {
    const int nrow = 1 + n_cosz, ldp = fold_ldp(n_etrue), ldg = fold_ld(8 * ((max_ereco + 7) / 8));
    const size_t over = (size_t)(fold_rows(nrow) - nrow) * (ldg > ldp ? ldg : ldp), after = grad_scratch_doubles(n_etrue, max_ereco);
    return sizeof(double) * ((size_t)4 * nrow * ldp + (size_t)2 * nrow * ldg + after + (over > after ? over - after : 0));
}
constexpr int kMCgrad = 4;   // row tiles the gradient folds keep in flight (4 accumulators each: the register budget of two CTAs per SM)

__device__ __forceinline__ double rcp_pos(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    return fma(fma(-x, r, 1.0), r, r);
}

// Poisson term and its derivative with respect to the raw expectation (likelihood_statistical.jl:16-24):
//   e' = (e + 1e-9 + sqrt((e - 1e-9)^2 + 1e-18)) / 2,   d e'/d e = (1 + (e - 1e-9) / sqrt(...)) / 2,
//   d term / d e' = 1 - m / e'  (m > 0)  or  1  (m = 0).
__device__ __forceinline__ double poisson_term_grad(double raw_e, double m, double logm, uint32_t& fl, double& dterm) {
    if (raw_e < 0.0) fl |= (raw_e > -1e-3) ? SOLAROSC_FLAG_NEG_WARN : SOLAROSC_FLAG_NEG_ERROR;
    const double d = raw_e - 1e-9;
    const double v = fma(d, d, 1e-18);
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
    double g = v * r, h = 0.5 * r;
    double e = fma(-h, g, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    e = fma(-h, g, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    g = fma(fma(-g, g, v), h, g);                       // sqrt(v); h = 1 / (2 sqrt(v))
    const double ep = 0.5 * (raw_e + 1e-9 + g);
    const double dep = fma(d, h, 0.5);
    const double t = ep - m + m * (logm - log_pos(ep));
    dterm = ((m > 0.0) ? fma(-m, rcp_pos(ep), 1.0) : 1.0) * dep;
    return (m > 0.0) ? t : ep;
}

// Sums v[0..N) over the block; the totals land in out[0..N) (shared memory) after the closing barrier.
template <int N>
__device__ __forceinline__ void block_reduce(const double (&v)[N], double* red /* [kWarps][N] */, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();   // red / out of an earlier reduction have been consumed
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double s = warp_sum(v[k]);
        if (lane == 0) red[warp * N + k] = s;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double s = 0.0;
        for (int w = 0; w < kWarps; ++w) s += red[w * N + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

// One CTA per parameter point.  osc holds the oscillation probabilities [n_points][2 proc][1 + n_cosz][n_etrue] as
// osc_kernel writes them; pbar_out receives d lnL / d P in the same layout (adjoint_kernel's input).  The gradient row
// is zeroed, then the flux, cross-section and background columns are written; the oscillation columns stay zero here.
__global__ void __launch_bounds__(kThreads, 2)
grad_kernel(const DevTables T, const GradLaunch L, const double* __restrict__ params, const double* __restrict__ osc,
            double* __restrict__ pbar_out, double* __restrict__ out_lnl, double* __restrict__ out_grad /* [n_points][n_params] */,
            uint32_t* __restrict__ out_flags) {
    extern __shared__ __align__(16) double smem[];
    __shared__ uint32_t redf[kWarps];
    const int nT = T.n_etrue, nZ = T.n_cosz, nrow = 1 + nZ, ldp = fold_ldp(nT), rows = nrow;   // no padding rows (grad_smem_bytes)
    const int kt = fold_kt(nT), mt = (nrow + 7) / 8;
    const int ntm = (L.max_ereco + 7) / 8, ldg = fold_ld(8 * ntm);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, fg = lane >> 2, ft = lane & 3;
    double* p8s = smem;                                       // probabilities, A operand of the forward fold
    double* phs = p8s + (size_t)rows * ldp;
    double* q8s = phs + (size_t)rows * ldp;                   // d lnL / d P, same layout
    double* qhs = q8s + (size_t)rows * ldp;
    double* g_es = qhs + (size_t)rows * ldp;                  // d lnL / d x, A operand of the backward fold
    double* g_cc = g_es + (size_t)rows * ldg;
    double* scr = g_cc + (size_t)rows * ldg;
    double* shape = scr;                  // [4 kt] s_i = norm exp(tilt l + curv l^2)
    double* eunit = shape + 4 * kt;       // [4 kt] exp(tilt l + curv l^2)
    double* hs = eunit + 4 * kt;          // [4 kt] H_i of the cross-section derivatives
    double* colg = hs + 4 * kt;           // [4 kt] scratch (unused tail keeps the layout simple)
    double* hbg_es = colg + 4 * kt;       // [8 ntm] BG_ES / 2 per reco bin
    double* hbg_cc = hbg_es + 8 * ntm;    // [8 ntm]
    double* cg_es = hbg_cc + 8 * ntm;     // [8 ntm] column sums of G (shifted bins)
    double* cg_cc = cg_es + 8 * ntm;      // [8 ntm]
    double* red = cg_cc + 8 * ntm;        // [kWarps][max(kGradShared, kGradBg)]
    double* tot = red + (size_t)kWarps * (kGradShared > kGradBg ? kGradShared : kGradBg);   // [kGradShared + kGradBg]

    const size_t b = blockIdx.x;
    const double* p = params + b * L.param_stride;
    const size_t var_stride = (size_t)2 * nrow * nT;
    const double* base = osc + b * var_stride;
    const int P = T.n_params;
    const double f8 = p[3], fh = p[4];

    for (int k = threadIdx.x; k < P; k += blockDim.x) out_grad[b * P + k] = 0.0;
    fold_zero_padding(p8s, nT, rows, ldp);
    {   // stage the base probabilities, clear the adjoint tile
        const int dq = (int)blockDim.x / nT, dr = (int)blockDim.x - dq * nT;
        int row = (int)threadIdx.x / nT, i = (int)threadIdx.x - row * nT;
        while (row < nrow) {
            p8s[row * ldp + i] = base[row * nT + i];
            phs[row * ldp + i] = base[(nrow + row) * nT + i];
            i += dr; row += dq;
            if (i >= nT) { i -= nT; ++row; }
        }
        for (int t = threadIdx.x; t < 2 * rows * ldp; t += blockDim.x) q8s[t] = 0.0;
    }
    bool oob = false;
    if (T.earth_unc)
        for (int k = 0; k < T.n_layers; ++k) {
            const double x = p[SOLAROSC_N_FIXED_PARAMS + k];
            oob |= !(x >= 0.0 && x <= 2.0);
        }

    double sh[kGradShared];   // 0 lnL-sum, 1 F8, 2 Fh, 3 norm, 4 tilt, 5 curv, 6.. directions
#pragma unroll
    for (int k = 0; k < kGradShared; ++k) sh[k] = 0.0;
    uint32_t fl = 0;
    const uint32_t tile_bytes = 8u * (uint32_t)ldp * 8u, gtile_bytes = 8u * (uint32_t)ldg * 8u;

    for (int di = 0; di < T.n_det; ++di) {
        const DevDetector& D = T.det[di];
        const double* pd = p + D.param_off;
        const int bte = D.nte_thr, btc = D.ntc_thr;
        const int ie0 = D.index_es - 1, ic0 = D.index_cc - 1;
        __syncthreads();
        // ---- per-point vectors ---------------------------------------------------------------------------
        for (int i = threadIdx.x; i < 4 * kt; i += blockDim.x) {
            double eu = 0.0;
            if (i < nT) {
                const double l = D.log_e_norm[i];
                eu = exp(fma(p[6], l, p[7] * (l * l)));
            }
            eunit[i] = eu; shape[i] = eu * p[5]; hs[i] = 0.0;
        }
        for (int j = threadIdx.x; j < 16 * ntm; j += blockDim.x) {
            double acc = 0.0;
            if (j < 8 * ntm) {
                if (j < D.n_es)
                    for (int q = 0; q < D.n_bg_es; ++q) {
                        const double raw = D.raw_bg_es[q * D.n_es + j];
                        acc += (D.bg_es_area[q] != 0.0) ? raw * (D.bg_es_area[q] * pd[0]) : raw;
                    }
                hbg_es[j] = 0.5 * acc;
            } else {
                const int jc = j - 8 * ntm;
                if (jc < D.n_cc)
                    for (int q = 0; q < D.n_bg_cc; ++q) {
                        const double raw = D.raw_bg_cc[q * D.n_cc + jc];
                        acc += (D.bg_cc_norm[q] >= 0) ? raw * (pd[D.n_norm_es + D.bg_cc_norm[q]] / 2.2e-6) : raw;
                    }
                hbg_cc[jc] = 0.5 * acc;
            }
        }
        __syncthreads();
        // ---- forward fold, Poisson terms, G = d lnL / d x ---------------------------------------------------
        for (int task = warp; task < bte + btc; task += kWarps) {
            const bool is_es = task < bte;
            if (is_es ? !D.es_mode : !D.cc_mode) continue;
            const int n0 = 8 * (is_es ? task : task - bte);
            const int ldb = 8 * (is_es ? bte : btc);
            const int nR = is_es ? D.n_es : D.n_cc;
            const int j0 = (is_es ? ie0 : ic0) + n0 + 2 * ft;
            const bool v0 = j0 < nR, v1 = j0 + 1 < nR;
            const double* m8 = (is_es ? D.a8_thr : D.b8_thr) + ft * ldb + n0 + fg;
            const double* mh = (is_es ? D.ah_thr : D.bh_thr) + ft * ldb + n0 + fg;
            // ES: constant parts C8, Ch ride in the accumulators; CC: none
            const double c80 = (is_es && v0) ? D.c8[j0] : 0.0, c81 = (is_es && v1) ? D.c8[j0 + 1] : 0.0;
            const double ch0 = (is_es && v0) ? D.ch[j0] : 0.0, ch1 = (is_es && v1) ? D.ch[j0 + 1] : 0.0;
            const double hb0 = v0 ? (is_es ? hbg_es : hbg_cc)[j0] : 0.0, hb1 = v1 ? (is_es ? hbg_es : hbg_cc)[j0 + 1] : 0.0;
            const double2* ol = is_es ? D.ol_es : D.ol_cc;
            double* gt = (is_es ? g_es : g_cc) + n0 + 2 * ft;
            const size_t bstep = (size_t)4 * ldb;
            for (int m0 = 0; m0 < mt; m0 += kMCgrad) {
                double x8[kMCgrad][2], xh[kMCgrad][2];
                uint32_t a8[kMCgrad];
#pragma unroll
                for (int c = 0; c < kMCgrad; ++c) {
                    x8[c][0] = c80; x8[c][1] = c81; xh[c][0] = ch0; xh[c][1] = ch1;
                    a8[c] = smem_addr(p8s + fg * ldp + ft) + (uint32_t)min(m0 + c, mt - 1) * tile_bytes;
                }
                const uint32_t dproc = smem_addr(phs) - smem_addr(p8s);
                const double *q8 = m8, *qh = mh;
                for (int ks = 0; ks < kt; ++ks) {
                    const double s = is_es ? 1.0 : shape[4 * ks + ft];
                    const double b8v = __ldg(q8) * s, bhv = __ldg(qh) * s;
                    q8 += bstep; qh += bstep;
#pragma unroll
                    for (int c = 0; c < kMCgrad; ++c) {
                        dmma_8x8x4(x8[c][0], x8[c][1], lds_f64(a8[c] + 32u * ks), b8v);
                        dmma_8x8x4(xh[c][0], xh[c][1], lds_f64(a8[c] + dproc + 32u * ks), bhv);
                    }
                }
#pragma unroll
                for (int c = 0; c < kMCgrad; ++c) {
                    const int row = 8 * (m0 + c) + fg;
                    const bool vr = m0 + c < mt && row < nrow;
                    const double w = (row == 0) ? 1.0 : T.wz[min(row, nrow - 1) - 1];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const bool ok = vr && (hh ? v1 : v0);
                        const double x = fma(f8, x8[c][hh], fma(fh, xh[c][hh], hh ? hb1 : hb0));
                        const double2 o = ok ? __ldg(ol + row * nR + j0 + hh) : make_double2(0.0, 0.0);
                        double dt;
                        const double t = poisson_term_grad(ok ? w * x : 1.0, o.x, o.y, fl, dt);
                        const double g = ok ? -dt * w : 0.0;          // d lnL / d x
                        if (ok) {
                            sh[0] += t;
                            sh[1] = fma(g, x8[c][hh], sh[1]);
                            sh[2] = fma(g, xh[c][hh], sh[2]);
                        }
                        if (vr) gt[(size_t)row * ldg + hh] = g;      // (rows past the tile pair belong to the next array)
                    }
                }
            }
        }
        __syncthreads();
        // ---- background norms: column sums of G against the raw components (propagation_bg.jl:36-83) -----
        double bg[kGradBg];
#pragma unroll
        for (int k = 0; k < kGradBg; ++k) bg[k] = 0.0;
        for (int t = threadIdx.x; t < 8 * (bte + btc); t += blockDim.x) {
            const bool is_es = t < 8 * bte;
            if (is_es ? !D.es_mode : !D.cc_mode) continue;
            const int jj = is_es ? t : t - 8 * bte, j = (is_es ? ie0 : ic0) + jj;
            const int nR = is_es ? D.n_es : D.n_cc;
            if (j >= nR) continue;
            const double* gcol = (is_es ? g_es : g_cc) + jj;
            double cs = 0.0;
            for (int row = 0; row < nrow; ++row) cs += gcol[(size_t)row * ldg];
            if (is_es) {
                double a = 0.0;
                for (int q = 0; q < D.n_bg_es; ++q) a = fma(D.raw_bg_es[q * D.n_es + j], D.bg_es_area[q], a);
                bg[0] = fma(cs, 0.5 * a, bg[0]);
            } else {
#pragma unroll
                for (int k = 0; k < SOLAROSC_MAX_BG; ++k) {
                    double a = 0.0;
                    for (int q = 0; q < D.n_bg_cc; ++q)
                        if (D.bg_cc_norm[q] == k) a += D.raw_bg_cc[q * D.n_cc + j];
                    bg[1 + k] = fma(cs, 0.5 * a / 2.2e-6, bg[1 + k]);
                }
            }
        }
        // ---- backward fold: Pbar = F (G_ES A^T + diag(s) G_CC B^T); H_i for the cross-section shape ---------
        const int nb = (nT + 7) / 8, kp = 4 * kt;
        for (int task = warp; task < nb; task += kWarps) {
            const int n0 = 8 * task;
            const int icol = min(n0 + fg, kp - 1);                  // row of the fold matrices this lane's B element sits in
            const int i0 = n0 + 2 * ft;                             // this lane's two true-energy bins
            for (int m0 = 0; m0 < mt; m0 += kMCgrad) {
                double a8v[kMCgrad][2], ahv[kMCgrad][2];
                uint32_t ga[kMCgrad];
#pragma unroll
                for (int c = 0; c < kMCgrad; ++c) {
                    a8v[c][0] = a8v[c][1] = ahv[c][0] = ahv[c][1] = 0.0;
                    ga[c] = (uint32_t)min(m0 + c, mt - 1) * gtile_bytes + (uint32_t)(fg * ldg + ft) * 8u;
                }
                if (D.cc_mode) {
                    const double* q8 = D.b8_thr + (size_t)icol * 8 * btc + ft;
                    const double* qh = D.bh_thr + (size_t)icol * 8 * btc + ft;
                    const uint32_t gbase = smem_addr(g_cc);
                    for (int ks = 0; ks < 2 * btc; ++ks) {
                        const double b8v = __ldg(q8 + 4 * ks), bhv = __ldg(qh + 4 * ks);
#pragma unroll
                        for (int c = 0; c < kMCgrad; ++c) {
                            const double gv = lds_f64(gbase + ga[c] + 32u * ks);
                            dmma_8x8x4(a8v[c][0], a8v[c][1], gv, b8v);
                            dmma_8x8x4(ahv[c][0], ahv[c][1], gv, bhv);
                        }
                    }
                    // cross-section shape: H_i = sum_rows F8 P8 (G_CC B8^T) + Fh Ph (G_CC Bh^T)
                    double h0 = 0.0, h1 = 0.0;
#pragma unroll
                    for (int c = 0; c < kMCgrad; ++c) {
                        const int row = 8 * (m0 + c) + fg;
                        if (m0 + c < mt && row < nrow) {
                            if (i0 < nT) h0 += f8 * p8s[row * ldp + i0] * a8v[c][0] + fh * phs[row * ldp + i0] * ahv[c][0];
                            if (i0 + 1 < nT) h1 += f8 * p8s[row * ldp + i0 + 1] * a8v[c][1] + fh * phs[row * ldp + i0 + 1] * ahv[c][1];
                        }
                    }
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) {
                        h0 += __shfl_xor_sync(0xffffffffu, h0, o);
                        h1 += __shfl_xor_sync(0xffffffffu, h1, o);
                    }
                    if (fg == 0) {
                        if (i0 < kp) hs[i0] += h0;
                        if (i0 + 1 < kp) hs[i0 + 1] += h1;
                    }
                    const double s0 = shape[min(i0, kp - 1)], s1 = shape[min(i0 + 1, kp - 1)];
#pragma unroll
                    for (int c = 0; c < kMCgrad; ++c) { a8v[c][0] *= s0; a8v[c][1] *= s1; ahv[c][0] *= s0; ahv[c][1] *= s1; }
                }
                if (D.es_mode) {
                    const double* q8 = D.a8_thr + (size_t)icol * 8 * bte + ft;
                    const double* qh = D.ah_thr + (size_t)icol * 8 * bte + ft;
                    const uint32_t gbase = smem_addr(g_es);
                    for (int ks = 0; ks < 2 * bte; ++ks) {
                        const double b8v = __ldg(q8 + 4 * ks), bhv = __ldg(qh + 4 * ks);
#pragma unroll
                        for (int c = 0; c < kMCgrad; ++c) {
                            const double gv = lds_f64(gbase + ga[c] + 32u * ks);
                            dmma_8x8x4(a8v[c][0], a8v[c][1], gv, b8v);
                            dmma_8x8x4(ahv[c][0], ahv[c][1], gv, bhv);
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < kMCgrad; ++c) {
                    const int row = 8 * (m0 + c) + fg;
                    if (m0 + c < mt && row < nrow) {
                        if (i0 < nT) { q8s[row * ldp + i0] += f8 * a8v[c][0]; qhs[row * ldp + i0] += fh * ahv[c][0]; }
                        if (i0 + 1 < nT) { q8s[row * ldp + i0 + 1] += f8 * a8v[c][1]; qhs[row * ldp + i0 + 1] += fh * ahv[c][1]; }
                    }
                }
            }
        }
        __syncthreads();
        // ---- cross-section parameters (propagation_osc.jl:185-196): s_i = norm e_i, e_i = exp(tilt l + curv l^2) --
        if (D.cc_mode)
            for (int i = threadIdx.x; i < nT; i += blockDim.x) {
                const double l = D.log_e_norm[i], hv = hs[i];
                sh[3] = fma(hv, eunit[i], sh[3]);
                sh[4] = fma(hv * shape[i], l, sh[4]);
                sh[5] = fma(hv * shape[i], l * l, sh[5]);
            }
        block_reduce(bg, red, tot + kGradShared);
        if (threadIdx.x == 0) {
            if (D.n_norm_es > 0) out_grad[b * P + D.param_off] = tot[kGradShared];
            for (int k = 0; k < D.n_norm_cc; ++k) out_grad[b * P + D.param_off + D.n_norm_es + k] = tot[kGradShared + 1 + k];
        }
    }
    // ---- d lnL / d P to HBM: adjoint_kernel carries it through the oscillation stage ------------------------------------
    __syncthreads();
    {
        double* dst = pbar_out + b * var_stride;
        for (int e = threadIdx.x; e < 2 * nrow * nT; e += blockDim.x) {
            const int proc = e / (nrow * nT), rem = e - proc * nrow * nT, row = rem / nT, i = rem - row * nT;
            dst[e] = (proc ? qhs : q8s)[row * ldp + i];
        }
    }
    for (int o = 16; o > 0; o >>= 1) fl |= __shfl_xor_sync(0xffffffffu, fl, o);
    if (lane == 0) redf[warp] = fl;
    block_reduce(sh, red, tot);
    if (threadIdx.x == 0) {
        uint32_t f = 0;
        for (int w = 0; w < kWarps; ++w) f |= redf[w];
        double lnl = -tot[0];
        if (f & SOLAROSC_FLAG_NEG_ERROR) lnl = NAN;
        if (oob) { lnl = -INFINITY; f = SOLAROSC_FLAG_EARTH_OOB; }
        else if (lnl != lnl) f |= (f & SOLAROSC_FLAG_NEG_ERROR) ? 0u : SOLAROSC_FLAG_NONFINITE;
        out_lnl[b] = lnl;
        if (out_flags) out_flags[b] = f;
    }
    if (oob) {   // a constant -Inf has no slope (the reference's dual numbers carry zero partials there)
        __syncthreads();
        for (int k = threadIdx.x; k < P; k += blockDim.x) out_grad[b * P + k] = 0.0;
        return;
    }
    if (threadIdx.x >= 1 && threadIdx.x < kGradShared)
        out_grad[b * P + threadIdx.x + 2] = tot[threadIdx.x];   // fluxes -> columns 3, 4; cross-section norm / tilt / curvature -> 5..7
}

}  // namespace solarosc
