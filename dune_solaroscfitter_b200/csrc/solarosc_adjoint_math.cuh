// solarosc_adjoint_math.cuh -- exact derivatives of the oscillation stage (K1 + K2) by an adjoint sweep.
//
// The reference differentiates total_llh with ForwardDiff (mcmc.jl:102-103): dual numbers through the eigen-systems,
// the layer products and the MSW formula.  grad_kernel (solarosc_grad.cuh) already carries the likelihood's sensitivity
// back through the Poisson sum and the response fold to  Pbar[proc][row][i] = d lnL / d P_proc[row][i]  on the FP64
// tensor pipe.  This file carries it the rest of the way, exactly, in one backward sweep per (path, energy):
//
//   night[p,e] = Pday(e) + K(e) (P1e[p,e] - P0),  K = c13^2 cos 2theta_m,  P1e = |psi_1|^2,
//   psi_1 = e_1^T U_N ... U_1 u,   U_k = exp(-i t_k S_k),  S_k = diag(0, dm21, dm31) + a_k u u^T   (solarosc_device.cuh)
//
//   d psi_1 = sum_k beta_k^dagger dU_k phi_{k-1} + beta_0^dagger du,   phi_k = U_k phi_{k-1},  beta_{k-1} = U_k^dagger beta_k,
//   dU = Q (Gamma o (Q^T dS Q)) Q^T,  Gamma_ij = (e_i - e_j) / (mu_i - mu_j),  Gamma_ii = -i t e_i,  e_j = exp(-i t mu_j)
//
// (Daleckii-Krein).  The forward states are not stored: U_k is unitary, so the backward sweep un-propagates phi
// next to the adjoint beta, which starts as c e_1 with c = 2 w K psi_1 -- then Re(beta^dagger x) is the weighted
// derivative of |psi_1|^2 directly.  In the eigenbasis of step k (pa = Q^T phi_k, bb = Q^T beta_k, aa = conj(E) pa)
//   pair (i,j):  Y_ij += Re[(e_i - e_j) (conj(bb_i) aa_j + conj(bb_j) aa_i)]
//   diagonal i:  Z_i  += len_k Im[conj(bb_i) pa_i]                       (i = 1, 2: the common phase mu_0 t drops out)
// are accumulated per (layer, energy) over all paths -- five numbers that do not depend on the direction -- and the
// directions come in only at the end, through D = Q^T dS Q of each table entry:
//   d/d dm21: D_ij = q_i[1] q_j[1];   d/d earth_norm_l: D_ij = (a/n_l)(q_i.u)(q_j.u) on layer l only;
//   d/d sin2_th12, sin2_th13: D_ij = a ((q_i.du)(q_j.u) + (q_i.u)(q_j.du)), plus the initial state: Re(beta_0).du
//   g += sum_{i<j} Y_ij D_ij / (mu_i - mu_j) + (c/E) sum_{i=1,2} Z_i (D_ii - D_00).
// The MSW factors Pday, K and P0 are differentiated analytically (osc.jl:77-90,185-207).  One backward step costs
// about 160 FP64 instructions next to the 64 of the forward step: all nine oscillation-stage derivatives for the
// price of ~3.5 forward stages, where central differences took 18.
#pragma once
#include "solarosc_device.cuh"

namespace solarosc {

// Eigen table entry of the adjoint kernel: all three eigenvectors.
//   [0..2] q_0  [3..5] q_1  [6..8] q_2  [9] c (mu_1-mu_0)/E  [10] c (mu_2-mu_0)/E  [11..13] q_0.u, q_1.u, q_2.u
// 14 doubles = 112 bytes: 16-byte aligned, and the pitch keeps the eight lanes of a quarter-warp on distinct banks.
constexpr int kEntryA = 14;

SOLAROSC_HD void make_entry_adj(double dm21, double a, double e_gev, double u1, double u2, double u3, double* ent) {
    double q[9], w2, w3;
    eig_rank_one(dm21, kDm31, a, u1, u2, u3, q, w2, w3);
    const double f = kPhaseQT / e_gev;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ent[3 * i + 0] = q[0 + i]; ent[3 * i + 1] = q[3 + i]; ent[3 * i + 2] = q[6 + i];
        ent[11 + i] = fma(q[6 + i], u3, fma(q[3 + i], u2, q[0 + i] * u1));
    }
    ent[9] = f * w2; ent[10] = f * w3;
}

// forward step on the full state (every step: the backward sweep starts from the complete final state)
SOLAROSC_HD void adj_step_fwd(const double* e, double len, State& p) {
    double s1, c1, s2, c2;
    sincos_qt(len, e[9], s1, c1);
    sincos_qt(len, e[10], s2, c2);
    const double w1r = fma(e[5], p.cr, fma(e[4], p.br, e[3] * p.ar)), w1i = fma(e[5], p.ci, fma(e[4], p.bi, e[3] * p.ai));
    const double w2r = fma(e[8], p.cr, fma(e[7], p.br, e[6] * p.ar)), w2i = fma(e[8], p.ci, fma(e[7], p.bi, e[6] * p.ai));
    const double t1r = fma(c1, w1r, fma(s1, w1i, -w1r)), t1i = fma(c1, w1i, fma(-s1, w1r, -w1i));
    const double t2r = fma(c2, w2r, fma(s2, w2i, -w2r)), t2i = fma(c2, w2i, fma(-s2, w2r, -w2i));
    p.ar = fma(e[6], t2r, fma(e[3], t1r, p.ar)); p.ai = fma(e[6], t2i, fma(e[3], t1i, p.ai));
    p.br = fma(e[7], t2r, fma(e[4], t1r, p.br)); p.bi = fma(e[7], t2i, fma(e[4], t1i, p.bi));
    p.cr = fma(e[8], t2r, fma(e[5], t1r, p.cr)); p.ci = fma(e[8], t2i, fma(e[5], t1i, p.ci));
}

struct AdjAcc {
    double y01, y02, y12, z1, z2;
};

// One backward step: phi = phi_k -> phi_{k-1}, beta = beta_k -> beta_{k-1}; returns this step's five accumulands.
SOLAROSC_HD AdjAcc adj_step_back(const double* e, double len, State& phi, State& beta) {
    double s1, c1, s2, c2;
    sincos_qt(len, e[9], s1, c1);     // e_j = c_j - i s_j
    sincos_qt(len, e[10], s2, c2);
    // eigenbasis components of phi_k and beta_k
    const double p0r = fma(e[2], phi.cr, fma(e[1], phi.br, e[0] * phi.ar)), p0i = fma(e[2], phi.ci, fma(e[1], phi.bi, e[0] * phi.ai));
    const double p1r = fma(e[5], phi.cr, fma(e[4], phi.br, e[3] * phi.ar)), p1i = fma(e[5], phi.ci, fma(e[4], phi.bi, e[3] * phi.ai));
    const double p2r = fma(e[8], phi.cr, fma(e[7], phi.br, e[6] * phi.ar)), p2i = fma(e[8], phi.ci, fma(e[7], phi.bi, e[6] * phi.ai));
    const double b0r = fma(e[2], beta.cr, fma(e[1], beta.br, e[0] * beta.ar)), b0i = fma(e[2], beta.ci, fma(e[1], beta.bi, e[0] * beta.ai));
    const double b1r = fma(e[5], beta.cr, fma(e[4], beta.br, e[3] * beta.ar)), b1i = fma(e[5], beta.ci, fma(e[4], beta.bi, e[3] * beta.ai));
    const double b2r = fma(e[8], beta.cr, fma(e[7], beta.br, e[6] * beta.ar)), b2i = fma(e[8], beta.ci, fma(e[7], beta.bi, e[6] * beta.ai));
    // aa_j = conj(e_j) pa_j: the components of phi_{k-1};  cb_j = conj(e_j) bb_j: those of beta_{k-1}   (aa_0 = pa_0, cb_0 = bb_0)
    const double a1r = fma(c1, p1r, -s1 * p1i), a1i = fma(c1, p1i, s1 * p1r);
    const double a2r = fma(c2, p2r, -s2 * p2i), a2i = fma(c2, p2i, s2 * p2r);
    const double d1r = fma(c1, b1r, -s1 * b1i), d1i = fma(c1, b1i, s1 * b1r);
    const double d2r = fma(c2, b2r, -s2 * b2i), d2i = fma(c2, b2i, s2 * b2r);
    AdjAcc o;
    {   // pairs: v = conj(bb_i) aa_j + conj(bb_j) aa_i,  Y = Re[(e_i - e_j) v]
        // (0,1): e_0 - e_1 = (1 - c1) + i s1
        const double vr = fma(b0r, a1r, fma(b0i, a1i, fma(b1r, p0r, b1i * p0i)));
        const double vi = fma(b0r, a1i, fma(-b0i, a1r, fma(b1r, p0i, -b1i * p0r)));
        o.y01 = fma(1.0 - c1, vr, -s1 * vi);
    }
    {   // (0,2)
        const double vr = fma(b0r, a2r, fma(b0i, a2i, fma(b2r, p0r, b2i * p0i)));
        const double vi = fma(b0r, a2i, fma(-b0i, a2r, fma(b2r, p0i, -b2i * p0r)));
        o.y02 = fma(1.0 - c2, vr, -s2 * vi);
    }
    {   // (1,2): e_1 - e_2 = (c1 - c2) - i (s1 - s2)
        const double vr = fma(b1r, a2r, fma(b1i, a2i, fma(b2r, a1r, b2i * a1i)));
        const double vi = fma(b1r, a2i, fma(-b1i, a2r, fma(b2r, a1i, -b2i * a1r)));
        o.y12 = fma(c1 - c2, vr, (s1 - s2) * vi);
    }
    o.z1 = len * fma(b1r, p1i, -b1i * p1r);   // len Im[conj(bb_1) pa_1]
    o.z2 = len * fma(b2r, p2i, -b2i * p2r);
    // un-propagate: x_{k-1} = x_k + sum_{n=1,2} q_n (conj(e_n) - 1) (q_n . x_k)
    const double f1r = a1r - p1r, f1i = a1i - p1i, f2r = a2r - p2r, f2i = a2i - p2i;
    phi.ar = fma(e[6], f2r, fma(e[3], f1r, phi.ar)); phi.ai = fma(e[6], f2i, fma(e[3], f1i, phi.ai));
    phi.br = fma(e[7], f2r, fma(e[4], f1r, phi.br)); phi.bi = fma(e[7], f2i, fma(e[4], f1i, phi.bi));
    phi.cr = fma(e[8], f2r, fma(e[5], f1r, phi.cr)); phi.ci = fma(e[8], f2i, fma(e[5], f1i, phi.ci));
    const double g1r = d1r - b1r, g1i = d1i - b1i, g2r = d2r - b2r, g2i = d2i - b2i;
    beta.ar = fma(e[6], g2r, fma(e[3], g1r, beta.ar)); beta.ai = fma(e[6], g2i, fma(e[3], g1i, beta.ai));
    beta.br = fma(e[7], g2r, fma(e[4], g1r, beta.br)); beta.bi = fma(e[7], g2i, fma(e[4], g1i, beta.bi));
    beta.cr = fma(e[8], g2r, fma(e[5], g1r, beta.cr)); beta.ci = fma(e[8], g2i, fma(e[5], g1i, beta.ci));
    return o;
}

// Mixing-angle derivatives of the initial state u = (c12 c13, s12 c13, s13) with respect to x = sin2_th12 and
// y = sin2_th13 (get_mixing_parameters clamps both to [0, 1], propagation_osc.jl:30-36: flat outside, and the square
// roots are singular on the boundary itself -- ForwardDiff returns Inf/NaN there; zero is returned here).
struct MixDeriv {
    double du_x[3], du_y[3];
};
SOLAROSC_HD MixDeriv mix_deriv(double s12sq_raw, double s13sq_raw) {
    MixDeriv m{};
    const double x = fmin(fmax(s12sq_raw, 0.0), 1.0), y = fmin(fmax(s13sq_raw, 0.0), 1.0);
    const double c12 = sqrt(1.0 - x), s12 = sqrt(x), c13 = sqrt(1.0 - y), s13 = sqrt(y);
    if (s12sq_raw > 0.0 && s12sq_raw < 1.0) { m.du_x[0] = -0.5 * c13 / c12; m.du_x[1] = 0.5 * c13 / s12; m.du_x[2] = 0.0; }
    if (s13sq_raw > 0.0 && s13sq_raw < 1.0) { m.du_y[0] = -0.5 * c12 / c13; m.du_y[1] = -0.5 * s12 / c13; m.du_y[2] = 0.5 / s13; }
    return m;
}

// Direction contractions of one table entry with its accumulated (Y, Z): adds to g[0] (sin2_th12), g[1] (sin2_th13),
// g[2] (dm2_21) and returns the earth-norm derivative of the entry's own layer.  a = norm * lookup * E, a_over_n = lookup * E.
SOLAROSC_HD double adj_entry_directions(const double* e, const AdjAcc& s, double a, double a_over_n, double e_gev, const MixDeriv& md,
                                        double* g) {
    // eigenvalue gaps from the tabulated phase rates: k_j = kPhaseQT (mu_j - mu_0) / E
    const double inv = e_gev / kPhaseQT;
    const double w1 = e[9] * inv, w2 = e[10] * inv;                 // mu_1 - mu_0, mu_2 - mu_0
    auto safe = [](double num, double den) { return den != 0.0 ? num / den : 0.0; };
    const double r01 = safe(s.y01, -w1), r02 = safe(s.y02, -w2), r12 = safe(s.y12, w1 - w2);
    const double zt = kPhase / e_gev, z1 = zt * s.z1, z2 = zt * s.z2;
    // contraction with a symmetric D: sum_{i<j} r_ij D_ij + z1 (D_11 - D_00) + z2 (D_22 - D_00)
    auto contract = [&](double d00, double d11, double d22, double d01, double d02, double d12) {
        return fma(r01, d01, fma(r02, d02, fma(r12, d12, fma(z1, d11 - d00, z2 * (d22 - d00)))));
    };
    const double qu0 = e[11], qu1 = e[12], qu2 = e[13];
    // dm2_21: dS = diag(0, 1, 0)
    g[2] += contract(e[1] * e[1], e[4] * e[4], e[7] * e[7], e[1] * e[4], e[1] * e[7], e[4] * e[7]);
    // mixing angles: dS = a (du u^T + u du^T)
    {
        const double v0 = fma(e[2], md.du_x[2], fma(e[1], md.du_x[1], e[0] * md.du_x[0]));
        const double v1 = fma(e[5], md.du_x[2], fma(e[4], md.du_x[1], e[3] * md.du_x[0]));
        const double v2 = fma(e[8], md.du_x[2], fma(e[7], md.du_x[1], e[6] * md.du_x[0]));
        g[0] += a * contract(2.0 * v0 * qu0, 2.0 * v1 * qu1, 2.0 * v2 * qu2, fma(v0, qu1, v1 * qu0), fma(v0, qu2, v2 * qu0), fma(v1, qu2, v2 * qu1));
    }
    {
        const double v0 = fma(e[2], md.du_y[2], fma(e[1], md.du_y[1], e[0] * md.du_y[0]));
        const double v1 = fma(e[5], md.du_y[2], fma(e[4], md.du_y[1], e[3] * md.du_y[0]));
        const double v2 = fma(e[8], md.du_y[2], fma(e[7], md.du_y[1], e[6] * md.du_y[0]));
        g[1] += a * contract(2.0 * v0 * qu0, 2.0 * v1 * qu1, 2.0 * v2 * qu2, fma(v0, qu1, v1 * qu0), fma(v0, qu2, v2 * qu0), fma(v1, qu2, v2 * qu1));
    }
    // earth norm of this layer: dS = (a / n) u u^T
    return a_over_n * contract(qu0 * qu0, qu1 * qu1, qu2 * qu2, qu0 * qu1, qu0 * qu2, qu1 * qu2);
}

// MSW day probability, K = c13^2 cos 2theta_m and their derivatives with respect to (sin2_th12, sin2_th13, dm2_21)
// at one density (osc.jl:77-90): v = {Pday, K}, d[0..2] = dPday, d[3..5] = dK.
SOLAROSC_HD void msw_deriv(double e, double ne, double dm21, double x12, double y, bool x_live, bool y_live, double* v, double* d) {
    const double c2 = 1.0 - 2.0 * x12, c13sq = 1.0 - y;
    const double a0e = 2.0 * 1.4142135623730951 * ne * kGF * e;
    const double beta = a0e * c13sq / dm21;
    const double X = c2 - beta, s2 = 1.0 - c2 * c2;
    const double R2 = fma(X, X, s2), R = sqrt(R2), c2m = X / R, iR3 = 1.0 / (R2 * R);
    const double dX = s2 * iR3;                       // d c2m / dX at fixed c2
    const double dc_x = x_live ? -2.0 * (dX + X * c2 * iR3) : 0.0;
    const double dc_y = y_live ? dX * (a0e / dm21) : 0.0;
    const double dc_m = dX * (beta / dm21);
    v[0] = fma(0.5 * c13sq * c13sq, fma(c2, c2m, 1.0), y * y);
    v[1] = c13sq * c2m;
    const double h = 0.5 * c13sq * c13sq;
    d[0] = x_live ? h * fma(c2, dc_x, -2.0 * c2m) : 0.0;
    d[1] = y_live ? fma(h * c2, dc_y, fma(-c13sq, fma(c2, c2m, 1.0), 2.0 * y)) : 0.0;
    d[2] = h * c2 * dc_m;
    d[3] = c13sq * dc_x;
    d[4] = y_live ? fma(c13sq, dc_y, -c2m) : 0.0;
    d[5] = c13sq * dc_m;
}

}  // namespace solarosc
