// solarosc_device.cuh -- device-side tables and FP64 math helpers for libsolarosc_b200.
//
// Everything here is written for sm_100a (B200).  The hot path is FP64-pipe bound (SURVEY.md
// section 8(d)): no tensor-core shaped work exists in it, so the design rules that matter are
// warp-uniform control flow, conflict-free shared-memory tables and as few FP64 instructions per
// Earth-layer step as the algebra allows.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

#include "../../include/solarosc_b200.h"

namespace solarosc {

constexpr double kGF = 5.4489e-5;                 // reference osc.jl:51
constexpr double kPhase = 2.5338653580781976;     // reference osc.jl:361
constexpr double kPhaseQT = kPhase * 0.63661977236758134308;   // the same rate in quarter turns (x 2/pi)
constexpr double kDm31 = 2.5e-3;                  // reference osc.jl:44 (theta23, deltaCP drop out of P1e)
constexpr int kMaxDet = 4;
constexpr int kMaxFoldTasks = 48;
constexpr int kMaxCos = 64;

// one applied layer step: 0-based layer and length [km]; 16 bytes so a step's metadata is one load
struct __align__(16) Seg {
    double len;
    int layer, pad;
};

struct DevDetector {
    int es_mode, cc_mode, n_es, n_cc, index_es, index_cc, n_bg_es, n_bg_cc, n_norm_es, n_norm_cc, param_off;
    const double* log_e_norm;             // [n_etrue]
    // fold matrices built at create (see fold_stage): [4*ceil(n_etrue/4)][8*ceil(n_ereco/8)], zero padded
    const double *a8, *ah;                // ES, 8B / hep
    const double *b8, *bh;                // CC, 8B / hep (without the per-point cross-section shape)
    const double *c8, *ch;                // ES additive constants per reco bin, [8*ceil(n_es/8)]
    // the same matrices with the columns below the likelihood threshold removed: column j' is reco bin index_xx - 1 + j'
    const double *a8_thr, *ah_thr, *b8_thr, *bh_thr;
    int nte_thr, ntc_thr;                 // their widths in 8-column blocks
    // likelihood fold: warp that runs work item t (fold_stage; longest-first assignment made at create);
    // fold_ntask = 0 when the list does not fit, the kernel then deals the items round-robin
    int fold_ntask;
    unsigned char fold_owner[kMaxFoldTasks];
    const double *raw_bg_es, *raw_bg_cc;  // [n_bg][n_ereco]
    double bg_es_area[SOLAROSC_MAX_BG];   // face area per ES component (0 => not scaled by a nuisance)
    int bg_cc_norm[SOLAROSC_MAX_BG];      // index of the CC nuisance scaling component i, -1 => fixed
    // data stored [1+n_cosz][n_ereco] (row 0 = day) and log(data) alongside
    const double *obs_es, *obs_cc, *logobs_es, *logobs_cc;
    const double2 *ol_es, *ol_cc;         // (m, log m) interleaved, same indexing
    // ---- inclusive / semi-inclusive / angular branches (solarosc_general.cuh); all zero for the exclusive analysis
    int incl, ang, n_cos;                 // 0 / 1 inclusive / 2 semi-inclusive; angular ES; cos-scatter bins
    int cc_out;                           // CC arrays are computed and enter the likelihood (cc_mode && incl != 1)
    double f_below;                       // CC_split.below on the CC backgrounds (1 unless semi-inclusive)
    const double *xs8, *xsh;              // CC projected on the ES bins (x f_above), layout of a8
    const double *yp8, *yph, *yc8, *ych;  // mis-identified ES in the CC bins: matrices (layout of b8) and constants
    const double *angm, *bgang;           // [n_es][n_cos]: mask_c ES.angular[c][j], BG.angular[c][j]
    const double *ang_sa, *ang_sb;        // [n_es]: their sums over the cos-scatter bins
    // cos-scatter bins with a non-zero table entry somewhere (the others -- the hemisphere the cos cut removes -- have an
    // expectation of exactly 0 whatever the parameters: their Poisson terms depend on the data only and are summed once,
    // at set_data, into llh_dead_const)
    int n_cos_live;
    unsigned char cos_live[kMaxCos];
    double llh_dead_const;
    const double2* ol_es_ang;             // angular ES data (m, log m), [1 + n_cosz][n_es][n_cos]
};

struct DevTables {
    int n_etrue, n_e, n_cosz, n_sub, n_paths, n_layers, n_solar, n_seg, n_det, n_params;
    int n_ltbl;                // layers some path actually crosses (1 + the largest layer id in use): the eigen table
                               // holds only these (the default binning never reaches PREM's innermost layer)
    int fast, earth_unc;
    int night_only;            // likelihoods sum the night rows only (solarosc_config.night_only)
    const double* e_calc;      // [n_e]
    const Seg* seg;            // [n_seg] the APPLIED steps of all paths, path after path (pad = 1 on the first step of a mirrored path)
    const double2* seg_mir;    // [n_seg] (dr, drh) per step: phase-rotation constants of mirrored paths
    const int2* path_se;       // [n_cosz*n_sub] (first, one-past-last) step of the paths of sorted row zr, sub-path fastest
    const int* row_order;      // [n_cosz] coarse zenith bin of sorted row zr (longest paths first)
    const unsigned* row_layers;// [n_cosz] bit k set when some path of sorted row zr crosses layer k
    const double* earth_lookup;// [n_layers]
    const double* wz;          // [n_cosz]
    double avg_ne[2];
    const double* solar_ne;    // [n_solar]
    const double* solar_w[2];  // production fractions 8B / hep
    double solar_wsum[2];
    DevDetector det[kMaxDet];
};

// ------------------------------------------------------------------------------------------
// sin/cos of a phase that is known to be moderate (|x| < 2^30): two-constant Cody-Waite
// reduction done with FMAs (x - k*pi/2 is exact in the first FMA because k*PIO2_HI needs at most
// 53+31 bits and the difference is < 1 in magnitude on the grid of ulp(PIO2_HI)), then the
// classic fdlibm minimax kernels on [-pi/4, pi/4].  About 22 FP64 instructions against ~40 for
// CUDA's sincos(), which also carries a Payne-Hanek slow path these phases can never take.
// The largest phase of the default configuration is c*L*dm31/E ~ 7e4 rad.
// ------------------------------------------------------------------------------------------
// Coefficients of the reduction and of the fdlibm sin/cos minimax kernels.  On the device they live
// in constant memory so that every DFMA takes its coefficient as a constant-bank operand (literal
// doubles are otherwise rebuilt with two UMOVs per use).
#define SOLAROSC_TRIG_TABLE                                                                           \
    6.36619772367581382433e-01, 1.57079632679489655800e+00, 6.12323399573676603587e-17,               \
        1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,          \
        -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,         \
        -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,         \
        2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02, 0.0
// Quarter-turn kernels of sincos_qt (tools/fit_quarter_turn_trig.py, Remez in 60-digit arithmetic): for |f| <= 1/2
//   sin(pi/2 f) = f (S0 + z S1 + ... + z^5 S5),  cos(pi/2 f) = 1 + z (C0 + z C1 + ... + z^5 C5),  z = f^2,
// absolute errors 3.4e-15 and 3.3e-16 evaluated in double -- three orders of magnitude below the 1e-11 rad
// rounding noise that phases of up to 7e4 rad carry anyway.
#define SOLAROSC_QT_TABLE                                                                             \
    1.57079632679488989666e+00, -6.45964097504311052234e-01, 7.96926261558089099601e-02,              \
        -4.68175259216174061150e-03, 1.60429268258354377910e-04, -3.55639139782485082505e-06,         \
        -1.23370055013616908290e+00, 2.53669507900830593083e-01, -2.08634807532030522315e-02,         \
        9.19260101511700715264e-04, -2.52007040774374478794e-05, 4.66321043146625407711e-07
static __constant__ double kTrigDev[28] = {SOLAROSC_TRIG_TABLE, SOLAROSC_QT_TABLE};
static const double kTrigHost[28] = {SOLAROSC_TRIG_TABLE, SOLAROSC_QT_TABLE};

#if defined(__CUDA_ARCH__)
#define SOLAROSC_HD __host__ __device__ __forceinline__
#define SOLAROSC_TRIG(i) kTrigDev[i]
// 1/sqrt(x) and 1/x of a positive normal double: hardware seed (22 bits) and two Newton steps, no special cases --
// about a third of the instructions of rsqrt() / a division, and straight-line code that the scheduler can
// interleave with its neighbours.  Within 1-2 ulp.
__device__ __forceinline__ double rsq(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = fma(y, fma(-hx * y, y, 0.5), y);
    return fma(y, fma(-hx * y, y, 0.5), y);
}
__device__ __forceinline__ double rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(fma(-x, y, 1.0), y, y);
    return fma(fma(-x, y, 1.0), y, y);
}
__device__ __forceinline__ int lo_int(double x) { return __double2loint(x); }
// x with its sign flipped when `flip` has bit 31 set: an integer XOR, not an FP64-pipe negation
__device__ __forceinline__ double flip_sign(double x, int flip) {
    return __hiloint2double(__double2hiint(x) ^ (flip & 0x80000000), __double2loint(x));
}
#else
// The same helpers compile for the host so that tests can run the exact kernel arithmetic on the
// CPU (tests/csrc/host_math_check.cu); the library itself never executes them there.
#define SOLAROSC_HD __host__ __device__ inline
#define SOLAROSC_TRIG(i) kTrigHost[i]
inline double rsq(double x) { return 1.0 / std::sqrt(x); }
inline double rcp(double x) { return 1.0 / x; }
inline int lo_int(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(uint32_t)(b & 0xffffffff); }
inline double flip_sign(double x, int flip) { return (flip & 0x80000000) ? -x : x; }
#endif

SOLAROSC_HD void sincos_phase(double x, double& s, double& c) {
    const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: forces round-to-nearest-integer
    double kd = fma(x, SOLAROSC_TRIG(0), kMagic);
    const int q = lo_int(kd);
    kd -= kMagic;
    double r = fma(-kd, SOLAROSC_TRIG(1), x);
    r = fma(-kd, SOLAROSC_TRIG(2), r);
    const double z = r * r;
    // sin kernel
    double ps = fma(z, SOLAROSC_TRIG(3), SOLAROSC_TRIG(4));
    ps = fma(z, ps, SOLAROSC_TRIG(5));
    ps = fma(z, ps, SOLAROSC_TRIG(6));
    ps = fma(z, ps, SOLAROSC_TRIG(7));
    ps = fma(z, ps, SOLAROSC_TRIG(8));
    const double sr = fma(z * r, ps, r);
    // cos kernel
    double pc = fma(z, SOLAROSC_TRIG(9), SOLAROSC_TRIG(10));
    pc = fma(z, pc, SOLAROSC_TRIG(11));
    pc = fma(z, pc, SOLAROSC_TRIG(12));
    pc = fma(z, pc, SOLAROSC_TRIG(13));
    pc = fma(z, pc, SOLAROSC_TRIG(14));
    const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
    // quadrant q mod 4: 0 (sr, cr)   1 (cr, -sr)   2 (-sr, -cr)   3 (-cr, sr)
    const double ss = (q & 1) ? cr : sr;
    const double cc = (q & 1) ? sr : cr;
    s = flip_sign(ss, q << 30);          // bit 1 of q -> sign bit
    c = flip_sign(cc, (q + 1) << 30);
}

// sin/cos of the phase len * kq measured in QUARTER TURNS (kq = rate in quarter turns per km, tabulated per
// eigen entry): the product, its rounding to an integer number of quarter turns and the reduced argument all
// come out of two FMAs on the unrounded product -- f = len*kq - round(len*kq) is formed from the exact 106-bit
// product, so the reduction adds no error at all -- and the minimax kernels take f directly (the pi/2 lives in
// their coefficients).  16 FP64 instructions for phase product + sine + cosine, against 1 + 20 with
// sincos_phase.  |len * kq| < 2^31.
SOLAROSC_HD void sincos_qt(double len, double kq, double& s, double& c) {
    const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
    const double t = fma(len, kq, kMagic);
    const int q = lo_int(t);
    const double f = fma(len, kq, -(t - kMagic));
    const double z = f * f;
    double ps = fma(z, SOLAROSC_TRIG(21), SOLAROSC_TRIG(20));
    ps = fma(z, ps, SOLAROSC_TRIG(19));
    ps = fma(z, ps, SOLAROSC_TRIG(18));
    ps = fma(z, ps, SOLAROSC_TRIG(17));
    ps = fma(z, ps, SOLAROSC_TRIG(16));
    const double sr = f * ps;
    double pc = fma(z, SOLAROSC_TRIG(27), SOLAROSC_TRIG(26));
    pc = fma(z, pc, SOLAROSC_TRIG(25));
    pc = fma(z, pc, SOLAROSC_TRIG(24));
    pc = fma(z, pc, SOLAROSC_TRIG(23));
    pc = fma(z, pc, SOLAROSC_TRIG(22));
    const double cr = fma(z, pc, 1.0);
    // quadrant q mod 4: 0 (sr, cr)   1 (cr, -sr)   2 (-sr, -cr)   3 (-cr, sr)
    const double ss = (q & 1) ? cr : sr;
    const double cc = (q & 1) ? sr : cr;
    s = flip_sign(ss, q << 30);
    c = flip_sign(cc, (q + 1) << 30);
}

// ------------------------------------------------------------------------------------------
// Eigen-system of the matter Hamiltonian in the vacuum mass basis.
//
// The reference diagonalises H = U diag(0,dm21,dm31) U^dagger + rho*E * diag(0,-1,-1)
// (osc.jl:370-388, complex Hermitian 3x3).  U = U23 * D(delta) * O * D(delta)^dagger with
// O = U13(real) * U12 real orthogonal, and both U23 and D act trivially on the electron row, so
// H is unitarily equivalent, by a matter-independent rotation, to the REAL symmetric
//     S = diag(0, dm21, dm31) + a * u u^T  - a*I,     a = rho*E,  u = (c12 c13, s12 c13, s13),
// and P1e = |<nu_1| prod_k exp(-i S_k t_k) |u>|^2 does not depend on theta23 or deltaCP at all.
// S is a rank-one update of a diagonal matrix: eigenvalues from the trigonometric solution of the
// characteristic cubic, each polished by a Newton step in gap form (below); eigenvectors from
// v_i[j] ~ u_j * prod_{k != j} (mu_i - m_k) (no divisions, well defined when a or some u_j
// vanish).  Only eigenvalue differences are returned (the common phase drops out of |amplitude|^2).
// q[3*j+i] = component j of eigenvector i.
// ------------------------------------------------------------------------------------------
SOLAROSC_HD void eig_rank_one(double d21, double d31, double a, double u1, double u2, double u3,
                              double q[9], double& w2, double& w3) {
    const double m[3] = {0.0, d21, d31};
    const double u[3] = {u1, u2, u3};
    double s11 = a * u1 * u1, s22 = fma(a * u2, u2, d21), s33 = fma(a * u3, u3, d31);
    double s12 = a * u1 * u2, s13 = a * u1 * u3, s23 = a * u2 * u3;
    // Written without branches (selects only), so that two eigen-systems in flight interleave.
    double mu[3];
    const double p1 = s12 * s12 + s13 * s13 + s23 * s23;
    {
        const double tq = (s11 + s22 + s33) * (1.0 / 3.0);
        double b11 = s11 - tq, b22 = s22 - tq, b33 = s33 - tq;
        const double p2 = (b11 * b11 + b22 * b22 + b33 * b33 + 2.0 * p1) * (1.0 / 6.0);
        const double ip = rsq(p2), p = p2 * ip;      // p = sqrt(p2); p2 == 0 only on the diagonal shortcut below
        b11 *= ip; b22 *= ip; b33 *= ip;
        const double c12 = s12 * ip, c13 = s13 * ip, c23 = s23 * ip;
        const double det = b11 * (b22 * b33 - c23 * c23) - c12 * (c12 * b33 - c23 * c13) + c13 * (c12 * c23 - b22 * c13);
        const double r = fmin(1.0, fmax(-1.0, 0.5 * det));
        const double phi = acos(r) * (1.0 / 3.0);
        double sp, cp;
        sincos_phase(phi, sp, cp);
        // cos(phi + 2pi/3) = -cp/2 - sqrt(3)/2 * sp
        const double hi = fma(2.0 * p, cp, tq);
        const double lo = fma(2.0 * p, fma(-0.8660254037844386, sp, -0.5 * cp), tq);
        const bool diag = p1 == 0.0;                 // no coupling at all: the diagonal is the spectrum
        mu[2] = diag ? s33 : hi;
        mu[0] = diag ? s11 : lo;
        mu[1] = diag ? s22 : 3.0 * tq - lo - hi;
    }
    // ascending order (only needed on the diagonal shortcut, cheap)
    {
        double t = fmin(mu[0], mu[1]); mu[1] = fmax(mu[0], mu[1]); mu[0] = t;
        t = fmin(mu[1], mu[2]); mu[2] = fmax(mu[1], mu[2]); mu[1] = t;
        t = fmin(mu[0], mu[1]); mu[1] = fmax(mu[0], mu[1]); mu[0] = t;
    }
    // Root i sits between pole m_i and the next one (interlacing).  Its gaps to the poles, g_x = m_x - mu_i, are
    // polished with one Newton step on the characteristic polynomial written in those gaps,
    //     h(g_i) = g_i A + a u_i^2 B,   A = g_k g_l + a (u_k^2 g_l + u_l^2 g_k),   B = g_k g_l
    // (all three gaps move together with mu_i).  In gap form every product is formed from small, accurately known
    // factors, so the step converges quadratically to full RELATIVE accuracy of each gap from the cubic's roots, whose
    // error is absolute: when a u_i^2 is far below the rounding noise of the cubic (weak matter, tiny theta13) h is
    // linear in g_i and the step lands on g_i = -a u_i^2 B / A; when the matter term exceeds dm21 and two roots share
    // the neighbourhood of two close poles it still converges, where the fixed-point form of the same equation would
    // amplify the cubic's error by u_i^2 / u_k^2.  A decoupled state (u_i = 0) keeps its pole.
    double dl[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int k = (i == 0) ? 1 : 0, l = (i == 2) ? 1 : 2;
        double gi = m[i] - mu[i], gk = m[k] - mu[i], gl = m[l] - mu[i];
        const double ui2 = u[i] * u[i], uk2 = u[k] * u[k], ul2 = u[l] * u[l];
        const double B = gk * gl, Bp = gk + gl;
        const double A = fma(a, fma(uk2, gl, ul2 * gk), B), Ap = fma(a, uk2 + ul2, Bp);
        const double aui2 = a * ui2;
        const double h = fma(gi, A, aui2 * B), hp = fma(gi, Ap, fma(aui2, Bp, A));
        // (an exactly decoupled state sits exactly on its pole: g_i = 0, and its eigenvector is the unit vector below)
        const double dg = (aui2 == 0.0) ? gi : ((hp != 0.0) ? h * rcp(hp) : 0.0);
        gi -= dg; gk -= dg; gl -= dg;
        const double d = -gi;
        dl[i] = d;
        // eigenvector: v[j] ~ u_j prod_{x != j} (mu_i - m_x)
        const double vi = u[i] * (gk * gl), vk = -u[k] * (d * gl), vl = -u[l] * (d * gk);
        const double n2 = vi * vi + vk * vk + vl * vl;
        // an exactly decoupled state (u_i == 0), or everything underflowed, is the unit vector e_i
        const bool ok = n2 > 0.0 && n2 < 1e300;
        const double rn = rsq(ok ? n2 : 1.0);
        double e[3];
        e[i] = ok ? vi * rn : 1.0; e[k] = ok ? vk * rn : 0.0; e[l] = ok ? vl * rn : 0.0;
        q[0 + i] = e[0]; q[3 + i] = e[1]; q[6 + i] = e[2];
    }
    w2 = (m[1] - m[0]) + (dl[1] - dl[0]);
    w3 = (m[2] - m[0]) + (dl[2] - dl[0]);
}

// ------------------------------------------------------------------------------------------
// Eigen table entry of one (layer, fine energy): what the propagation loop needs and nothing else.
// With eigenvectors q_0, q_1, q_2 (q_0 belongs to the lowest eigenvalue) one constant-density step is
//     psi <- Q diag(1, e^{-i phi_1}, e^{-i phi_2}) Q^T psi = psi + sum_{n=1,2} q_n (e^{-i phi_n} - 1) (q_n . psi),
// so q_0 and the common phase are never needed (reference osc_kernel, osc.jl:360-364, applied to a
// vector instead of forming the 3x3 products of osc.jl:442).
//   [0..2] q_1   [3..5] q_2   [6] c*(mu_1-mu_0)/E   [7] c*(mu_2-mu_0)/E   [8] q_1.u   [9] q_2.u
// with c = 2.5338653580781976 * 2/pi: the phase rates are kept in quarter turns per km (sincos_qt).
// Ten doubles = 80 bytes: entries are 16-byte aligned, so a step fetches its eight numbers with four 128-bit
// loads, and the 80-byte pitch between the entries of consecutive energies (the lanes of a warp) keeps every
// quarter-warp of such a load on distinct banks.
// ------------------------------------------------------------------------------------------
constexpr int kEntry = 10;

SOLAROSC_HD void make_entry(double dm21, double a, double e_gev, double u1, double u2, double u3, double* ent) {
    double q[9], w2, w3;
    eig_rank_one(dm21, kDm31, a, u1, u2, u3, q, w2, w3);
    const double f = kPhaseQT / e_gev;   // phase = 2.5338653580781976 * (l/e) * lambda  (osc.jl:361), in quarter turns
    ent[0] = q[1]; ent[1] = q[4]; ent[2] = q[7];
    ent[3] = q[2]; ent[4] = q[5]; ent[5] = q[8];
    ent[6] = f * w2; ent[7] = f * w3;
    ent[8] = fma(q[7], u3, fma(q[4], u2, q[1] * u1));
    ent[9] = fma(q[8], u3, fma(q[5], u2, q[2] * u1));
}

struct State { double ar, ai, br, bi, cr, ci; };

// what a layer step needs of an entry, fetched with 128-bit loads
struct Ent {
    double q1x, q1y, q1z, q2x, q2y, q2z, k1, k2;
};
SOLAROSC_HD Ent load_ent(const double* e) {
    const double2* v = reinterpret_cast<const double2*>(e);
    const double2 a = v[0], b = v[1], c = v[2], d = v[3];
    return Ent{a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
}

// e^{-i phi_n} = c_n - i s_n for both excited eigenstates of this entry
SOLAROSC_HD void step_phases(const Ent& e, double len, double& s1, double& c1, double& s2, double& c2) {
    sincos_qt(len, e.k1, s1, c1);
    sincos_qt(len, e.k2, s2, c2);
}

// first step of a path: psi = u is real, q_n.u is tabulated; t_n = (e^{-i phi_n} - 1) (q_n.u)
SOLAROSC_HD void step_first(const double* ep, double len, double u1, double u2, double u3, State& p) {
    const Ent e = load_ent(ep);
    const double2 d = reinterpret_cast<const double2*>(ep)[4];
    double s1, c1, s2, c2;
    step_phases(e, len, s1, c1, s2, c2);
    const double t1r = fma(c1, d.x, -d.x), t1i = -s1 * d.x, t2r = fma(c2, d.y, -d.y), t2i = -s2 * d.y;
    p.ar = fma(e.q2x, t2r, fma(e.q1x, t1r, u1)); p.ai = fma(e.q2x, t2i, e.q1x * t1i);
    p.br = fma(e.q2y, t2r, fma(e.q1y, t1r, u2)); p.bi = fma(e.q2y, t2i, e.q1y * t1i);
    p.cr = fma(e.q2z, t2r, fma(e.q1z, t1r, u3)); p.ci = fma(e.q2z, t2i, e.q1z * t1i);
}

// the only step of a one-layer path: first and last at once, only the nu_1 component is formed
SOLAROSC_HD void step_only(const double* ep, double len, double u1, State& p) {
    const Ent e = load_ent(ep);
    const double2 d = reinterpret_cast<const double2*>(ep)[4];
    double s1, c1, s2, c2;
    step_phases(e, len, s1, c1, s2, c2);
    p.ar = fma(e.q2x, fma(c2, d.y, -d.y), fma(e.q1x, fma(c1, d.x, -d.x), u1));
    p.ai = -fma(e.q2x, s2 * d.y, e.q1x * (s1 * d.x));
}

// t_n = (e^{-i phi_n} - 1) (q_n . psi), formed as c w + s (-i w) - w without ever forming e^{-i phi} - 1
SOLAROSC_HD void step_project(const Ent& e, double len, const State& p, double& t1r, double& t1i, double& t2r, double& t2i) {
    double s1, c1, s2, c2;
    step_phases(e, len, s1, c1, s2, c2);
    const double w1r = fma(e.q1z, p.cr, fma(e.q1y, p.br, e.q1x * p.ar)), w1i = fma(e.q1z, p.ci, fma(e.q1y, p.bi, e.q1x * p.ai));
    const double w2r = fma(e.q2z, p.cr, fma(e.q2y, p.br, e.q2x * p.ar)), w2i = fma(e.q2z, p.ci, fma(e.q2y, p.bi, e.q2x * p.ai));
    t1r = fma(c1, w1r, fma(s1, w1i, -w1r)); t1i = fma(c1, w1i, fma(-s1, w1r, -w1i));
    t2r = fma(c2, w2r, fma(s2, w2i, -w2r)); t2i = fma(c2, w2i, fma(-s2, w2r, -w2i));
}

SOLAROSC_HD void step_mid(const double* ep, double len, State& p) {
    const Ent e = load_ent(ep);
    double t1r, t1i, t2r, t2i;
    step_project(e, len, p, t1r, t1i, t2r, t2i);
    p.ar = fma(e.q2x, t2r, fma(e.q1x, t1r, p.ar)); p.ai = fma(e.q2x, t2i, fma(e.q1x, t1i, p.ai));
    p.br = fma(e.q2y, t2r, fma(e.q1y, t1r, p.br)); p.bi = fma(e.q2y, t2i, fma(e.q1y, t1i, p.bi));
    p.cr = fma(e.q2z, t2r, fma(e.q1z, t1r, p.cr)); p.ci = fma(e.q2z, t2i, fma(e.q1z, t1i, p.ci));
}

// last step: only the nu_1 component of the state is read afterwards
SOLAROSC_HD void step_last(const double* ep, double len, State& p) {
    const Ent e = load_ent(ep);
    double t1r, t1i, t2r, t2i;
    step_project(e, len, p, t1r, t1i, t2r, t2i);
    p.ar = fma(e.q2x, t2r, fma(e.q1x, t1r, p.ar)); p.ai = fma(e.q2x, t2i, fma(e.q1x, t1i, p.ai));
}

// P(nu_1 -> nu_e) of one path at one energy.  tbl points at the entry of layer 0 for this energy,
// layer_stride is the distance in doubles between the entries of consecutive layers.
SOLAROSC_HD double path_p1e(const double* tbl, size_t layer_stride, const int* app_layer, const double* app_len, int s0,
                            int s1, double u1, double u2, double u3) {
    State p;
    const double* e = tbl + (size_t)app_layer[s0] * layer_stride;
    if (s1 - s0 == 1) {   // one layer: first and last step at once
        step_only(e, app_len[s0], u1, p);
    } else {
        step_first(e, app_len[s0], u1, u2, u3, p);
        for (int s = s0 + 1; s < s1 - 1; ++s) step_mid(tbl + (size_t)app_layer[s] * layer_stride, app_len[s], p);
        step_last(tbl + (size_t)app_layer[s1 - 1] * layer_stride, app_len[s1 - 1], p);
    }
    return fma(p.ar, p.ar, p.ai * p.ai);   // |<nu_1|X|nu_e>|^2 (osc.jl:444)
}

// ------------------------------------------------------------------------------------------
// Mirrored paths.  A chord through the layered Earth crosses layers k_0, k_1, .., k_h, .., k_1, k_0, and
// every step matrix M = exp(-i S t) is complex SYMMETRIC (S is real symmetric), so the amplitude of
// osc.jl:436-446,
//     <nu_1| M'_0 M'_1 .. M'_{h-1}  M_h  M_{h-1} .. M_1 M_0 |u>  =  chi^T (M_h psi),
//     psi = M_{h-1} .. M_0 u,      chi = M'_{h-1} .. M'_0 e_1        (a plain, un-conjugated dot product),
// can be formed by walking psi (from the nu_e end) and chi (from the nu_1 end) inward TOGETHER: both cross
// the same layer at the same energy, so they share the eigen entry, and the lengths of the two crossings of an
// interior layer differ only by the rounding of the path builder (<= 1e-6 km), so chi's phase factors follow
// from psi's by a second-order rotation through delta = (pi/2) dlen k (6 instructions per phase instead of a
// 16-instruction sincos_qt; delta^3 / 6 < 1e-15 is checked per path when the handle is created -- paths that
// are not palindromes, or whose mirrored lengths differ by more, take the generic loop).  The outermost pair
// (the detector's depth makes its two lengths differ by metres) takes two full phase evaluations.
// 108 FP64 instructions per interior pair of steps against 128, one table fetch instead of two.
// ------------------------------------------------------------------------------------------
// e^{-i (phi + delta)} from e^{-i phi}: delta = dr * k, dr = (pi/2)(len' - len), drh = -dr / 2 (both tabulated per step)
SOLAROSC_HD void phase_shift(double s, double c, double k, double dr, double drh, double& s2, double& c2) {
    const double d = dr * k, dh = drh * k;
    s2 = fma(fma(s, dh, c), d, s);    // s + c d - s d^2 / 2
    c2 = fma(fma(c, dh, -s), d, c);   // c - s d - c d^2 / 2
}

// psi <- psi + sum_n q_n (e^{-i phi_n} - 1)(q_n . psi) with the phase factors given
SOLAROSC_HD void step_apply(const Ent& e, double s1, double c1, double s2, double c2, State& p) {
    const double w1r = fma(e.q1z, p.cr, fma(e.q1y, p.br, e.q1x * p.ar)), w1i = fma(e.q1z, p.ci, fma(e.q1y, p.bi, e.q1x * p.ai));
    const double w2r = fma(e.q2z, p.cr, fma(e.q2y, p.br, e.q2x * p.ar)), w2i = fma(e.q2z, p.ci, fma(e.q2y, p.bi, e.q2x * p.ai));
    const double t1r = fma(c1, w1r, fma(s1, w1i, -w1r)), t1i = fma(c1, w1i, fma(-s1, w1r, -w1i));
    const double t2r = fma(c2, w2r, fma(s2, w2i, -w2r)), t2i = fma(c2, w2i, fma(-s2, w2r, -w2i));
    p.ar = fma(e.q2x, t2r, fma(e.q1x, t1r, p.ar)); p.ai = fma(e.q2x, t2i, fma(e.q1x, t1i, p.ai));
    p.br = fma(e.q2y, t2r, fma(e.q1y, t1r, p.br)); p.bi = fma(e.q2y, t2i, fma(e.q1y, t1i, p.bi));
    p.cr = fma(e.q2z, t2r, fma(e.q1z, t1r, p.cr)); p.ci = fma(e.q2z, t2i, fma(e.q1z, t1i, p.ci));
}

// outermost pair: psi starts from u over length len_p, chi from e_1 = (1, 0, 0) over length len_x (q_n . e_1 = q_nx)
SOLAROSC_HD void step_first_pair(const double* ep, double len_p, double len_x, double u1, double u2, double u3, State& p, State& x) {
    const Ent e = load_ent(ep);
    const double2 d = reinterpret_cast<const double2*>(ep)[4];
    double s1, c1, s2, c2;
    step_phases(e, len_p, s1, c1, s2, c2);
    {
        const double t1r = fma(c1, d.x, -d.x), t1i = -s1 * d.x, t2r = fma(c2, d.y, -d.y), t2i = -s2 * d.y;
        p.ar = fma(e.q2x, t2r, fma(e.q1x, t1r, u1)); p.ai = fma(e.q2x, t2i, e.q1x * t1i);
        p.br = fma(e.q2y, t2r, fma(e.q1y, t1r, u2)); p.bi = fma(e.q2y, t2i, e.q1y * t1i);
        p.cr = fma(e.q2z, t2r, fma(e.q1z, t1r, u3)); p.ci = fma(e.q2z, t2i, e.q1z * t1i);
    }
    step_phases(e, len_x, s1, c1, s2, c2);
    {
        const double t1r = fma(c1, e.q1x, -e.q1x), t1i = -s1 * e.q1x, t2r = fma(c2, e.q2x, -e.q2x), t2i = -s2 * e.q2x;
        x.ar = fma(e.q2x, t2r, fma(e.q1x, t1r, 1.0)); x.ai = fma(e.q2x, t2i, e.q1x * t1i);
        x.br = fma(e.q2y, t2r, e.q1y * t1r); x.bi = fma(e.q2y, t2i, e.q1y * t1i);
        x.cr = fma(e.q2z, t2r, e.q1z * t1r); x.ci = fma(e.q2z, t2i, e.q1z * t1i);
    }
}

// interior pair: both states cross the layer of entry ep, psi over len, chi over len + dlen
SOLAROSC_HD void step_pair(const double* ep, double len, double dr, double drh, State& p, State& x) {
    const Ent e = load_ent(ep);
    double s1, c1, s2, c2, s1x, c1x, s2x, c2x;
    step_phases(e, len, s1, c1, s2, c2);
    phase_shift(s1, c1, e.k1, dr, drh, s1x, c1x);
    phase_shift(s2, c2, e.k2, dr, drh, s2x, c2x);
    step_apply(e, s1, c1, s2, c2, p);
    step_apply(e, s1x, c1x, s2x, c2x, x);
}

// innermost layer: |chi^T (M psi)|^2
SOLAROSC_HD double step_middle(const double* ep, double len, State& p, const State& x) {
    step_mid(ep, len, p);
    const double re = fma(-x.ci, p.ci, fma(x.cr, p.cr, fma(-x.bi, p.bi, fma(x.br, p.br, fma(-x.ai, p.ai, x.ar * p.ar)))));
    const double im = fma(x.ci, p.cr, fma(x.cr, p.ci, fma(x.bi, p.br, fma(x.br, p.bi, fma(x.ai, p.ar, x.ar * p.ai)))));
    return fma(re, re, im * im);
}

// Whether the applied steps s0 .. s1-1 form a mirrored path the pair walk is exact for: an odd number (>= 3) of
// steps, palindromic layers, and interior mirrored lengths so close that the second-order phase rotation is exact
// to 1e-15 for every phase rate up to kq_max (quarter turns per km).  Host side (create).
inline bool mirror_eligible(const int* app_layer, const double* app_len, int s0, int s1, double kq_max) {
    const int n = s1 - s0;
    if (n < 3 || (n & 1) == 0) return false;
    for (int i = 0; i < n / 2; ++i) {
        if (app_layer[s0 + i] != app_layer[s1 - 1 - i]) return false;
        if (i == 0) continue;
        const double delta = 1.5707963267948966 * std::fabs(app_len[s1 - 1 - i] - app_len[s0 + i]) * kq_max;
        if (!(delta <= 1.8e-5)) return false;   // delta^3 / 6 <= 1e-15
    }
    return true;
}

// P1e of a mirrored path at one energy (mir[s] = (dr, drh) of applied step s, used for the interior steps of the first half)
SOLAROSC_HD double path_p1e_mirror(const double* tbl, size_t layer_stride, const int* app_layer, const double* app_len,
                                   const double* mir, int s0, int s1, double u1, double u2, double u3) {
    State p, x;
    const int h = (s1 - s0) / 2;
    step_first_pair(tbl + (size_t)app_layer[s0] * layer_stride, app_len[s0], app_len[s1 - 1], u1, u2, u3, p, x);
    for (int i = 1; i < h; ++i)
        step_pair(tbl + (size_t)app_layer[s0 + i] * layer_stride, app_len[s0 + i], mir[2 * (s0 + i)], mir[2 * (s0 + i) + 1], p, x);
    return step_middle(tbl + (size_t)app_layer[s0 + h] * layer_stride, app_len[s0 + h], p, x);
}

// MSW survival probability pieces (reference osc.jl:77-90): returns cos(2 theta12_matter)
SOLAROSC_HD double msw_c2m(double e, double ne, double dm21, double one_m_s13sq, double c2) {
    double acc = 2.0 * 1.4142135623730951 * ne * kGF * e * one_m_s13sq;
    double x = c2 - acc / dm21;
    return x * rsq(fma(x, x, 1.0 - c2 * c2));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace solarosc
