/*
 * solarosc_oracle.c -- CPU oracle (plain C) for the ES+CC day/night likelihood hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  A CPU restatement of the reference's algorithm, used as the
 * checker by tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
 * bench.py.  The product library (libsolarosc_b200.so) neither links nor calls it.
 *
 * PARITY PINNING: the reference is Julia (not runnable in this image) and its own tests pin only
 * the Poisson term and the earth-norm bound (tests/test_likelihood_correctness.jl:5-17,164-182);
 * those are checked in tests/test_oracle.py.  For the MSW probability, the Earth-matter P1e, the
 * fold and the background normalisation no reference vector exists: "parity unpinned" -- the
 * oracle is validated by construction, by agreement with the independent numpy twin
 * (oracle/oracle_np.py, which uses LAPACK's eigh where this file uses a Jacobi solver) and by
 * physics known answers.
 *
 * It shares only the table layout (include/solarosc_b200.h) with the product.  Row-major
 * (col_major == 0) tables only.  Citations are reference paths.
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/solarosc_b200.h"

typedef double complex cplx;

#define G_F 5.4489e-5                 /* osc.jl:51  */
#define PHASE_CONST 2.5338653580781976 /* osc.jl:361 */
#define DM2_31 2.5e-3                 /* osc.jl:44  */
#define DELTA_CP (-1.611)

/* ---- 3x3 complex helpers ------------------------------------------------------------------ */
static void mat_mul(const cplx a[9], const cplx b[9], cplx c[9]) {
    cplx t[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            cplx s = 0;
            for (int k = 0; k < 3; ++k) s += a[3 * i + k] * b[3 * k + j];
            t[3 * i + j] = s;
        }
    memcpy(c, t, sizeof t);
}
static void mat_adj(const cplx a[9], cplx c[9]) {
    cplx t[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) t[3 * i + j] = conj(a[3 * j + i]);
    memcpy(c, t, sizeof t);
}

/* Cyclic Jacobi for a 3x3 Hermitian matrix: eigenvalues ascending in lam, eigenvectors in the
 * columns of v.  Stands in for StaticArrays' eigen(Hermitian(...)) called at osc.jl:366-388; the
 * result enters only through V f(lambda) V^dagger, which does not depend on eigenvector phases. */
static void eigh3(const cplx h[9], cplx v[9], double lam[3]) {
    cplx a[9];
    memcpy(a, h, sizeof a);
    /* Hermitian() reads the upper triangle */
    a[3] = conj(a[1]); a[6] = conj(a[2]); a[7] = conj(a[5]);
    a[0] = creal(a[0]); a[4] = creal(a[4]); a[8] = creal(a[8]);
    for (int i = 0; i < 9; ++i) v[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = cabs(a[1]) + cabs(a[2]) + cabs(a[5]);
        double diag = fabs(creal(a[0])) + fabs(creal(a[4])) + fabs(creal(a[8]));
        if (off <= 1e-300 || off < 1e-22 * diag) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                cplx apq = a[3 * p + q];
                double r = cabs(apq);
                if (r == 0.0) continue;
                cplx ph = apq / r; /* e^{i alpha} */
                double app = creal(a[3 * p + p]), aqq = creal(a[3 * q + q]);
                double tau = (aqq - app) / (2.0 * r);
                double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(tau * tau + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                cplx j[9];
                for (int i = 0; i < 9; ++i) j[i] = (i % 4 == 0) ? 1.0 : 0.0;
                j[3 * p + p] = c;
                j[3 * p + q] = s;
                j[3 * q + p] = -s * conj(ph);
                j[3 * q + q] = c * conj(ph);
                cplx jd[9], tmp[9];
                mat_adj(j, jd);
                mat_mul(jd, a, tmp);
                mat_mul(tmp, j, a);
                mat_mul(v, j, v);
                a[3 * p + q] = 0; a[3 * q + p] = 0;
                a[3 * p + p] = creal(a[3 * p + p]);
                a[3 * q + q] = creal(a[3 * q + q]);
            }
    }
    lam[0] = creal(a[0]); lam[1] = creal(a[4]); lam[2] = creal(a[8]);
    for (int i = 0; i < 2; ++i) /* sort ascending, columns follow */
        for (int k = 0; k < 2 - i; ++k)
            if (lam[k] > lam[k + 1]) {
                double tl = lam[k]; lam[k] = lam[k + 1]; lam[k + 1] = tl;
                for (int rr = 0; rr < 3; ++rr) {
                    cplx tv = v[3 * rr + k]; v[3 * rr + k] = v[3 * rr + k + 1]; v[3 * rr + k + 1] = tv;
                }
            }
}

/* ---- oscillation physics --------------------------------------------------------------------- */
typedef struct { double dm21, th12, th13, dm31, th23, dcp; } osc_pars;

static double clamp01(double x) { return x < 0 ? 0 : (x > 1 ? 1 : x); }

/* propagation_osc.jl:30-36 + osc.jl:44 */
static osc_pars mixing(const double *p) {
    osc_pars m;
    m.dm21 = p[2];
    m.th12 = asin(sqrt(clamp01(p[0])));
    m.th13 = asin(sqrt(clamp01(p[1])));
    m.dm31 = DM2_31;
    m.th23 = asin(sqrt(0.5));
    m.dcp = DELTA_CP;
    return m;
}

/* osc.jl:57-64 (SMatrix literals are column-major) */
static void pmns(const osc_pars *m, cplx u[9]) {
    double c23 = cos(m->th23), s23 = sin(m->th23), c13 = cos(m->th13), s13 = sin(m->th13);
    double c12 = cos(m->th12), s12 = sin(m->th12);
    cplx ep = cexp(I * m->dcp), em = cexp(-I * m->dcp);
    cplx u1[9] = {1, 0, 0, 0, c23, s23, 0, -s23, c23};
    cplx u2[9] = {c13, 0, s13 * em, 0, 1, 0, -s13 * ep, 0, c13};
    cplx u3[9] = {c12, s12, 0, -s12, c12, 0, 0, 0, 1};
    cplx t[9];
    mat_mul(u1, u2, t);
    mat_mul(t, u3, u);
}

/* osc.jl:77-90 */
static void msw_prob(double e, const osc_pars *m, double ne, double *prob, double *c2m) {
    double s2 = sin(m->th13); s2 *= s2;
    double c2 = cos(2 * m->th12);
    double acc = 2 * sqrt(2.0) * ne * G_F * e * (1 - s2);
    double beta = acc / m->dm21;
    double x = c2 - beta;
    *c2m = x / sqrt(x * x + (1 - c2 * c2));
    double s13m = sqrt(s2);
    *prob = 1.0 / 2 * (1 - s2) * (1 - s13m * s13m) * (1 + c2 * *c2m) + s2 * s13m * s13m;
}

/* osc.jl:360-364: K = V diag(conj(exp(i c (l/e) lambda))) V^dagger */
static void osc_kernel(const cplx v[9], const double lam[3], double e, double l, cplx k[9]) {
    cplx t[9], vd[9];
    double f = PHASE_CONST * (l / e);
    for (int j = 0; j < 3; ++j) {
        cplx ph = conj(cexp(I * (f * lam[j])));
        for (int i = 0; i < 3; ++i) t[3 * i + j] = v[3 * i + j] * ph;
    }
    mat_adj(v, vd);
    mat_mul(t, vd, k);
}

/* osc.jl:448-477 with osc_reduce_from_lookup (:436-446); p1e[ip*nE + ie] */
static void osc_prob_earth(const solarosc_tables *t, const osc_pars *m, const double *lookup, double *p1e) {
    int nE = 2 * t->n_etrue, nP = t->n_cosz * t->n_cosz_sub, nL = t->n_layers;
    cplx u[9], ud[9], h0[9] = {0}, heff[9], tmp[9];
    pmns(m, u);
    mat_adj(u, ud);
    h0[4] = m->dm21; h0[8] = m->dm31;
    mat_mul(u, h0, tmp);
    mat_mul(tmp, ud, heff);
    for (int ie = 0; ie < nE; ++ie) {
        double e = t->e_calc[ie];
        cplx V[SOLAROSC_MAX_LAYERS][9];
        double L[SOLAROSC_MAX_LAYERS][3];
        for (int k = 0; k < nL; ++k) { /* get_H, osc.jl:370-388 */
            cplx h[9];
            memcpy(h, heff, sizeof h);
            double re = lookup[k] * e;
            h[0] += re;
            h[0] -= re; h[4] -= re; h[8] -= re;
            eigh3(h, V[k], L[k]);
        }
        for (int ip = 0; ip < nP; ++ip) {
            int s0 = t->seg_start[ip], n = t->seg_start[ip + 1] - s0;
            /* reversed arrays: ridx[j] = idx[n-1-j], rlen[j] = len[n-1-j]; step i uses
             * ridx[n-1-i] = idx[i] and rlen[i] = len[n-1-i] */
            cplx x[9], k[9];
            int lay = t->seg_layer[s0 + 0] - 1;
            osc_kernel(V[lay], L[lay], e, t->seg_length[s0 + n - 1], x);
            for (int i = 1; i < n; ++i) {
                lay = t->seg_layer[s0 + i] - 1;
                osc_kernel(V[lay], L[lay], e, t->seg_length[s0 + n - 1 - i], k);
                mat_mul(x, k, x);
            }
            cplx amp = 0;
            for (int j = 0; j < 3; ++j) amp += ud[j] * x[3 * j]; /* (U^dagger X)[1,1] */
            p1e[(size_t)ip * nE + ie] = creal(amp) * creal(amp) + cimag(amp) * cimag(amp);
        }
    }
}

/* Shared oscillation stage (propagation_main.jl:47-54, propagation_osc.jl:52-156): nu_e probabilities
 * at analysis resolution, out[proc][row][i], proc 0 = 8B, 1 = hep, row 0 = day, 1+z = night bin z. */
static void osc_probs(const solarosc_tables *t, const solarosc_config *c, const double *p, double *out,
                      double *p1e_out) {
    int nE = 2 * t->n_etrue, nP = t->n_cosz * t->n_cosz_sub, nT = t->n_etrue, nZ = t->n_cosz, ns = t->n_cosz_sub;
    osc_pars m = mixing(p);
    double lookup[SOLAROSC_MAX_LAYERS];
    for (int k = 0; k < t->n_layers; ++k)
        lookup[k] = c->earth_uncertainty ? p[SOLAROSC_N_FIXED_PARAMS + k] * t->earth_lookup[k] : t->earth_lookup[k];
    double *p1e = malloc(sizeof(double) * nP * nE);
    double *day = malloc(sizeof(double) * nE), *night = malloc(sizeof(double) * nP * nE);
    osc_prob_earth(t, &m, lookup, p1e);
    if (p1e_out) memcpy(p1e_out, p1e, sizeof(double) * nP * nE);
    double c13sq = cos(m.th13); c13sq *= c13sq;
    double p0 = cos(m.th12) * cos(m.th12) * c13sq;
    for (int proc = 0; proc < 2; ++proc) {
        if (c->fast) { /* osc.jl:185-207 */
            double ne = proc == 0 ? t->avg_ne_8b : t->avg_ne_hep;
            for (int ie = 0; ie < nE; ++ie) {
                double pd, c2m;
                msw_prob(t->e_calc[ie], &m, ne, &pd, &c2m);
                day[ie] = pd;
                for (int ip = 0; ip < nP; ++ip)
                    night[(size_t)ip * nE + ie] = pd + c13sq * c2m * (p1e[(size_t)ip * nE + ie] - p0);
            }
        } else { /* osc.jl:139-182 */
            const double *w = proc == 0 ? t->solar_prod_8b : t->solar_prod_hep;
            double tot = 0;
            for (int r = 0; r < t->n_solar; ++r) tot += w[r];
            memset(night, 0, sizeof(double) * nP * nE);
            for (int ie = 0; ie < nE; ++ie) day[ie] = 0;
            for (int r = 0; r < t->n_solar; ++r)
                for (int ie = 0; ie < nE; ++ie) {
                    double pd, c2m;
                    msw_prob(t->e_calc[ie], &m, t->solar_ne[r], &pd, &c2m);
                    day[ie] += pd * w[r];
                    for (int ip = 0; ip < nP; ++ip) {
                        double dp = c13sq * c2m * (p1e[(size_t)ip * nE + ie] - p0);
                        night[(size_t)ip * nE + ie] += w[r] * (pd + dp);
                    }
                }
            for (int ie = 0; ie < nE; ++ie) day[ie] /= tot;
            for (size_t i = 0; i < (size_t)nP * nE; ++i) night[i] /= tot;
        }
        double *o = out + (size_t)proc * (1 + nZ) * nT;
        for (int i = 0; i < nT; ++i) o[i] = (day[2 * i] + day[2 * i + 1]) / 2; /* block_average(.,2) */
        for (int z = 0; z < nZ; ++z)
            for (int i = 0; i < nT; ++i) { /* block_average(., (N_COSZ_SUB, 2)), column-major sum order */
                double s = 0;
                for (int ce = 0; ce < 2; ++ce)
                    for (int cz = 0; cz < ns; ++cz) s += night[(size_t)(z * ns + cz) * nE + 2 * i + ce];
                o[(size_t)(1 + z) * nT + i] = s / (ns * 2);
            }
    }
    free(p1e); free(day); free(night);
}

/* ---- rates ---------------------------------------------------------------------------------------- */
static int n_norms(const int32_t *par, int n) {
    int c = 0;
    for (int i = 0; i < n; ++i) c += par[i] != 0;
    return c;
}

static int det_param_offset(const solarosc_tables *t, const solarosc_config *c, int det) {
    int off = SOLAROSC_N_FIXED_PARAMS + (c->earth_uncertainty ? t->n_layers : 0);
    for (int d = 0; d < det; ++d)
        off += n_norms(t->det[d].bg_es_par, t->det[d].n_bg_es) + n_norms(t->det[d].bg_cc_par, t->det[d].n_bg_cc);
    return off;
}

int oracle_n_params(const solarosc_tables *t, const solarosc_config *c) { return det_param_offset(t, c, t->n_det); }

/* propagation_bg.jl:6-18 */
static double es_conversion(int side, double norm) {
    double area;
    if (side == 0) area = 2 * 12e2 * 64e2;
    else if (side == 1) area = 2 * 12e2 * 12e2;
    else area = 2 * 12e2 * 64e2 + 2 * 12e2 * 12e2;
    return area * norm;
}

/* fold of one channel term: acc[row][j] += 0.5*((sum_i S[row][i] R[i][j]) * eff[j])  (propagation_core.jl:73-79) */
static void fold_add(const double *S, const double *R, const double *eff, int nrow, int nT, int nR, double *acc) {
    for (int row = 0; row < nrow; ++row)
        for (int j = 0; j < nR; ++j) {
            double s = 0;
            for (int i = 0; i < nT; ++i) s += S[(size_t)row * nT + i] * R[(size_t)i * nR + j];
            acc[(size_t)row * nR + j] += 0.5 * (s * eff[j]);
        }
}

/* propagateSamples, exclusive non-angular branch (propagation_main.jl:56-157) for one detector.
 * es[row][j], cc[row][j] with row 0 = day, 1+z = night. */
static void rates_det(const solarosc_tables *t, const solarosc_config *c, int det, const double *p, const double *osc,
                      double *es, double *cc, double *bg_es, double *bg_cc) {
    const solarosc_detector *d = &t->det[det];
    int nT = t->n_etrue, nZ = t->n_cosz, nrow = 1 + nZ, nEs = d->n_ereco_es, nCc = d->n_ereco_cc;
    const double *pd = p + det_param_offset(t, c, det);
    int nes = n_norms(d->bg_es_par, d->n_bg_es);
    double f8 = p[3], fh = p[4];
    const double *P8 = osc, *Ph = osc + (size_t)nrow * nT;
    /* backgrounds (propagation_bg.jl:36-83): all ES components use norm 1 */
    for (int j = 0; j < nEs; ++j) bg_es[j] = 0;
    for (int i = 0; i < d->n_bg_es; ++i) {
        double f = d->bg_es_par[i] != 0 ? es_conversion(d->bg_es_sides[i], pd[0]) : 1.0;
        for (int j = 0; j < nEs; ++j)
            bg_es[j] += d->bg_es_par[i] != 0 ? d->raw_bg_es[(size_t)i * nEs + j] * f : d->raw_bg_es[(size_t)i * nEs + j];
    }
    for (int j = 0; j < nCc; ++j) bg_cc[j] = 0;
    for (int i = 0, k = 0; i < d->n_bg_cc; ++i) {
        if (d->bg_cc_par[i] != 0) {
            double f = pd[nes + k] / 2.2e-6;
            ++k;
            for (int j = 0; j < nCc; ++j) bg_cc[j] += d->raw_bg_cc[(size_t)i * nCc + j] * f;
        } else
            for (int j = 0; j < nCc; ++j) bg_cc[j] += d->raw_bg_cc[(size_t)i * nCc + j];
    }
    double *S = malloc(sizeof(double) * nrow * nT);
    memset(es, 0, sizeof(double) * nrow * nEs);
    memset(cc, 0, sizeof(double) * nrow * nCc);
    if (d->es_mode) { /* propagation_osc.jl:164-182, propagation_reco.jl:10-29 */
        for (int fl = 0; fl < 2; ++fl) {
            const double *u8 = fl == 0 ? d->unosc_es_nue_8b : d->unosc_es_nuother_8b;
            const double *uh = fl == 0 ? d->unosc_es_nue_hep : d->unosc_es_nuother_hep;
            for (int i = 0; i < nT; ++i) {
                double p8 = fl == 0 ? P8[i] : 1 - P8[i], ph = fl == 0 ? Ph[i] : 1 - Ph[i];
                S[i] = u8[i] * p8 * f8 + uh[i] * ph * fh;
            }
            for (int z = 0; z < nZ; ++z)
                for (int i = 0; i < nT; ++i) {
                    double p8 = P8[(size_t)(1 + z) * nT + i], ph = Ph[(size_t)(1 + z) * nT + i];
                    if (fl) { p8 = 1 - p8; ph = 1 - ph; }
                    double w = t->exposure_weights[z];
                    S[(size_t)(1 + z) * nT + i] = u8[i] * (f8 * p8 * w) + uh[i] * (fh * ph * w);
                }
            fold_add(S, fl == 0 ? d->resp_es_nue : d->resp_es_nuother, fl == 0 ? d->eff_es_nue : d->eff_es_nuother,
                     nrow, nT, nEs, es);
        }
        for (int j = 0; j < nEs; ++j) es[j] += 0.5 * bg_es[j];
        for (int z = 0; z < nZ; ++z)
            for (int j = 0; j < nEs; ++j) es[(size_t)(1 + z) * nEs + j] += 0.5 * (bg_es[j] * t->exposure_weights[z]);
    }
    if (d->cc_mode) { /* propagation_osc.jl:184-207, propagation_reco.jl:117-130 */
        for (int i = 0; i < nT; ++i) {
            double l = d->log_e_norm[i];
            double shape = exp(p[6] * l + p[7] * (l * l)) * p[5];
            double c8 = d->unosc_cc_8b[i] * shape, ch = d->unosc_cc_hep[i] * shape;
            S[i] = c8 * P8[i] * f8 + ch * Ph[i] * fh;
            for (int z = 0; z < nZ; ++z) {
                double w = t->exposure_weights[z];
                S[(size_t)(1 + z) * nT + i] =
                    c8 * (f8 * P8[(size_t)(1 + z) * nT + i] * w) + ch * (fh * Ph[(size_t)(1 + z) * nT + i] * w);
            }
        }
        fold_add(S, d->resp_cc, d->eff_cc, nrow, nT, nCc, cc);
        for (int j = 0; j < nCc; ++j) cc[j] += 0.5 * bg_cc[j];
        for (int z = 0; z < nZ; ++z)
            for (int j = 0; j < nCc; ++j) cc[(size_t)(1 + z) * nCc + j] += 0.5 * (bg_cc[j] * t->exposure_weights[z]);
    }
    free(S);
}

/* ---- likelihood ---------------------------------------------------------------------------------------- */
/* likelihood_statistical.jl:1-28; returns the (negated) sum, ORs flag bits into *flags */
double oracle_poisson(const double *e_in, const double *m_in, size_t n, uint32_t *flags) {
    double llh = 0;
    for (size_t i = 0; i < n; ++i) {
        double raw = e_in[i], m = m_in[i];
        if (raw < 0) {
            if (raw > -1e-3) { if (flags) *flags |= SOLAROSC_FLAG_NEG_WARN; }
            else if (flags) *flags |= SOLAROSC_FLAG_NEG_ERROR;
        }
        double e = 0.5 * (raw + 1e-9 + sqrt((raw - 1e-9) * (raw - 1e-9) + 1e-18));
        llh += m > 0 ? e - m + m * (log(m) - log(e)) : e;
    }
    return -llh;
}

static double llh_channel(const double *r, const double *od, const double *on, int nrow, int nR, int idx, uint32_t *fl,
                          int night_only) {
    int k = idx - 1;
    /* likelihood_builder.jl:5-20,129-144; night_only (solarosc_config, BASELINE config 4) drops the day vector */
    double s = night_only ? 0.0 : oracle_poisson(r + k, od + k, nR - k, fl);
    double night = 0;
    for (int row = 1; row < nrow; ++row)
        night += oracle_poisson(r + (size_t)row * nR + k, on + (size_t)(row - 1) * nR + k, nR - k, fl);
    return s + night;
}

static int earth_oob(const solarosc_tables *t, const solarosc_config *c, const double *p) {
    if (!c->earth_uncertainty) return 0;
    for (int k = 0; k < t->n_layers; ++k) {
        double x = p[SOLAROSC_N_FIXED_PARAMS + k];
        if (!(0 <= x && x <= 2)) return 1; /* likelihood_debug.jl:1-4 */
    }
    return 0;
}

/* total_llh (likelihood_main.jl:30-35): shared oscillation stage, sum over detectors */
static double llh_point(const solarosc_tables *t, const solarosc_config *c, const double *p, uint32_t *flag) {
    uint32_t fl = 0;
    if (earth_oob(t, c, p)) { if (flag) *flag = SOLAROSC_FLAG_EARTH_OOB; return -INFINITY; }
    int nT = t->n_etrue, nrow = 1 + t->n_cosz;
    double *osc = malloc(sizeof(double) * 2 * nrow * nT);
    osc_probs(t, c, p, osc, NULL);
    double tot = 0;
    for (int det = 0; det < t->n_det; ++det) {
        const solarosc_detector *d = &t->det[det];
        double *es = malloc(sizeof(double) * nrow * d->n_ereco_es), *cc = malloc(sizeof(double) * nrow * d->n_ereco_cc);
        double *be = malloc(sizeof(double) * d->n_ereco_es), *bc = malloc(sizeof(double) * d->n_ereco_cc);
        rates_det(t, c, det, p, osc, es, cc, be, bc);
        double l = 0.0;
        if (d->es_mode) l += llh_channel(es, d->obs_es_day, d->obs_es_night, nrow, d->n_ereco_es, d->index_es, &fl, c->night_only);
        if (d->cc_mode) l += llh_channel(cc, d->obs_cc_day, d->obs_cc_night, nrow, d->n_ereco_cc, d->index_cc, &fl, c->night_only);
        tot += l;
        free(es); free(cc); free(be); free(bc);
    }
    free(osc);
    if (fl & SOLAROSC_FLAG_NEG_ERROR) tot = NAN; /* the reference raises here */
    if (flag) *flag = fl;
    return tot;
}

/* ---- exported entry points (ctypes) ----------------------------------------------------------------- */
int oracle_eval_llh(const solarosc_tables *t, const solarosc_config *c, const double *params, size_t n, double *out,
                    uint32_t *flags, int n_threads) {
    int P = oracle_n_params(t, c);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (long b = 0; b < (long)n; ++b) {
        uint32_t fl = 0;
        out[b] = llh_point(t, c, params + (size_t)b * P, &fl);
        if (flags) flags[b] = fl;
    }
    (void)n_threads;
    return 0;
}

int oracle_eval_oscprobs(const solarosc_tables *t, const solarosc_config *c, const double *params, size_t n,
                         double *osc /* [n][2][1+nZ][nT] */, double *p1e /* [n][nP][nE] or NULL */) {
    int P = oracle_n_params(t, c), nT = t->n_etrue, nrow = 1 + t->n_cosz;
    size_t np1 = (size_t)t->n_cosz * t->n_cosz_sub * 2 * nT;
    for (size_t b = 0; b < n; ++b)
        osc_probs(t, c, params + b * P, osc + b * 2 * nrow * nT, p1e ? p1e + b * np1 : NULL);
    return 0;
}

int oracle_eval_rates(const solarosc_tables *t, const solarosc_config *c, int det, const double *params, size_t n,
                      double *es /* [n][1+nZ][nEs] */, double *cc, double *bg_es, double *bg_cc) {
    int P = oracle_n_params(t, c), nT = t->n_etrue, nrow = 1 + t->n_cosz;
    const solarosc_detector *d = &t->det[det];
    double *osc = malloc(sizeof(double) * 2 * nrow * nT);
    for (size_t b = 0; b < n; ++b) {
        osc_probs(t, c, params + b * P, osc, NULL);
        rates_det(t, c, det, params + b * P, osc, es + b * nrow * d->n_ereco_es, cc + b * nrow * d->n_ereco_cc,
                  bg_es + b * d->n_ereco_es, bg_cc + b * d->n_ereco_cc);
    }
    free(osc);
    return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
